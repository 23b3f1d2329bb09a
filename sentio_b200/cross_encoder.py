"""Cross-encoder weights container for K5 (BERT-style sequence classifier, MiniLM-L6 shape by default).

Blob layout consumed by ``sb_ce_load`` (all fp32, row-major, torch ``nn.Linear`` convention W[out, in]):

    word_embeddings [V,H] | position_embeddings [P,H] | token_type_embeddings [T,H] | emb_ln.gamma [H] | emb_ln.beta [H]
    for each layer:  Wq [H,H] bq [H] | Wk bk | Wv bv | Wo [H,H] bo [H] | ln1.gamma ln1.beta [H]
                     W1 [I,H] b1 [I] | W2 [H,I] b2 [H] | ln2.gamma ln2.beta [H]
    pooler.W [H,H] pooler.b [H] | classifier.w [H] classifier.b [1]

The reference has no local cross-encoder (README.md:63, src/core/rerankers/jina_reranker.py:139 posts to the Jina API);
BASELINE.json config 4 defines the model as a random-init MiniLM-L6-shaped BERT.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

MINILM_L6 = dict(vocab_size=30522, hidden=384, layers=6, heads=12, intermediate=1536, max_pos=512, type_vocab=2,
                 ln_eps=1e-12)


@dataclass
class CrossEncoderWeights:
    config: dict
    tensors: dict = field(default_factory=dict)  # name -> np.float32 array

    # ------------------------------------------------------------------ naming
    @staticmethod
    def tensor_order(cfg: dict) -> list[tuple[str, tuple[int, ...]]]:
        V, H, L, I, P, T = (cfg["vocab_size"], cfg["hidden"], cfg["layers"], cfg["intermediate"], cfg["max_pos"],
                            cfg.get("type_vocab", 2))
        order = [("word_emb", (V, H)), ("pos_emb", (P, H)), ("type_emb", (T, H)), ("emb_ln_g", (H,)), ("emb_ln_b", (H,))]
        for l in range(L):
            p = f"l{l}."
            order += [(p + "wq", (H, H)), (p + "bq", (H,)), (p + "wk", (H, H)), (p + "bk", (H,)), (p + "wv", (H, H)),
                      (p + "bv", (H,)), (p + "wo", (H, H)), (p + "bo", (H,)), (p + "ln1_g", (H,)), (p + "ln1_b", (H,)),
                      (p + "w1", (I, H)), (p + "b1", (I,)), (p + "w2", (H, I)), (p + "b2", (H,)), (p + "ln2_g", (H,)),
                      (p + "ln2_b", (H,))]
        order += [("pool_w", (H, H)), ("pool_b", (H,)), ("cls_w", (H,)), ("cls_b", (1,))]
        return order

    def blob(self) -> np.ndarray:
        parts = []
        for name, shape in self.tensor_order(self.config):
            t = np.asarray(self.tensors[name], dtype=np.float32)
            if t.shape != shape:
                raise ValueError(f"{name}: expected shape {shape}, got {t.shape}")
            parts.append(t.reshape(-1))
        return np.concatenate(parts)

    # ------------------------------------------------------------------ constructors
    @classmethod
    def random(cls, cfg: dict, seed: int = 0, std: float = 0.02) -> "CrossEncoderWeights":
        """BERT-style init: N(0, std) matrices / embeddings, zero biases, unit LayerNorm gains."""
        rng = np.random.default_rng(seed)
        tensors = {}
        for name, shape in cls.tensor_order(cfg):
            base = name.split(".")[-1]
            if base.endswith("_g"):
                t = np.ones(shape, dtype=np.float32)
            elif base.startswith("b") or base.endswith("_b"):
                t = np.zeros(shape, dtype=np.float32)
            else:
                t = (rng.standard_normal(shape, dtype=np.float32) * np.float32(std)).astype(np.float32)
            tensors[name] = t
        return cls(dict(cfg), tensors)

    @classmethod
    def random_minilm_l6(cls, seed: int = 0) -> "CrossEncoderWeights":
        return cls.random(MINILM_L6, seed)

    @classmethod
    def from_hf_state_dict(cls, sd: dict, cfg: dict) -> "CrossEncoderWeights":
        """Import from a HuggingFace ``BertForSequenceClassification.state_dict()`` (values: torch tensors / arrays)."""
        def g(key):
            v = sd[key]
            return np.asarray(v.detach().cpu().numpy() if hasattr(v, "detach") else v, dtype=np.float32)

        t = {"word_emb": g("bert.embeddings.word_embeddings.weight"),
             "pos_emb": g("bert.embeddings.position_embeddings.weight"),
             "type_emb": g("bert.embeddings.token_type_embeddings.weight"),
             "emb_ln_g": g("bert.embeddings.LayerNorm.weight"), "emb_ln_b": g("bert.embeddings.LayerNorm.bias")}
        for l in range(cfg["layers"]):
            hp, p = f"bert.encoder.layer.{l}.", f"l{l}."
            t[p + "wq"], t[p + "bq"] = g(hp + "attention.self.query.weight"), g(hp + "attention.self.query.bias")
            t[p + "wk"], t[p + "bk"] = g(hp + "attention.self.key.weight"), g(hp + "attention.self.key.bias")
            t[p + "wv"], t[p + "bv"] = g(hp + "attention.self.value.weight"), g(hp + "attention.self.value.bias")
            t[p + "wo"], t[p + "bo"] = g(hp + "attention.output.dense.weight"), g(hp + "attention.output.dense.bias")
            t[p + "ln1_g"], t[p + "ln1_b"] = g(hp + "attention.output.LayerNorm.weight"), g(hp + "attention.output.LayerNorm.bias")
            t[p + "w1"], t[p + "b1"] = g(hp + "intermediate.dense.weight"), g(hp + "intermediate.dense.bias")
            t[p + "w2"], t[p + "b2"] = g(hp + "output.dense.weight"), g(hp + "output.dense.bias")
            t[p + "ln2_g"], t[p + "ln2_b"] = g(hp + "output.LayerNorm.weight"), g(hp + "output.LayerNorm.bias")
        t["pool_w"], t["pool_b"] = g("bert.pooler.dense.weight"), g("bert.pooler.dense.bias")
        t["cls_w"], t["cls_b"] = g("classifier.weight").reshape(-1), g("classifier.bias").reshape(-1)
        return cls(dict(cfg), t)
