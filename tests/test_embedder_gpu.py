"""§8f row 1: the on-device embedder (sb_enc_*) vs the HuggingFace BertModel forward of the same weights
(final [CLS] state -> projection -> L2 norm).  Tolerance: the unit-norm embedding is within 1e-3 of the oracle's in L2
(= 1e-3 relative, the tolerance north_star states for the floating-point stages), cosine > 1 - 1e-6, every element within
3e-4; BaseEmbedder surface; embed -> dense search chain.  (An emulation of the device numerics -- fp16 GEMM operands,
packed-half GELU -- predicts 4e-4 in L2 and 8e-5 per element: scripts/gelu_fp16_study.py.)"""
import asyncio

import numpy as np
import pytest

from oracle import cross_encoder as ce_oracle
from sentio_b200 import synth
from sentio_b200.cross_encoder import MINILM_L6, CrossEncoderWeights
from sentio_b200.embedder import B200Embedder, tokenize_for_embedding

pytestmark = pytest.mark.gpu


def _texts(n):
    flat, off = synth.text_corpus_tokens(n, vocab=3000)
    t = synth.texts_from_tokens(flat, off)
    t[0] = ""
    t[1] = "w1 " * 300
    return t


@pytest.mark.parametrize("project", [True, False])
def test_embeddings_match_huggingface_oracle(engine, project):
    model = ce_oracle.hf_model(MINILM_L6, seed=2)
    w = CrossEncoderWeights.from_hf_state_dict(model.state_dict(), MINILM_L6)
    rng = np.random.default_rng(9)
    pw = (rng.standard_normal((1024, 384)) / np.sqrt(384)).astype(np.float32) if project else None
    pb = (rng.standard_normal(1024) * 0.01).astype(np.float32) if project else None
    engine.enc_load(w.blob(), MINILM_L6, pw, pb)
    assert engine.enc_dim() == (1024 if project else 384)
    texts = _texts(40)
    ids, tt, lens = tokenize_for_embedding(texts, 128)
    got = engine.enc_embed(ids, tt, lens)
    want = ce_oracle.embed_from_cls(ce_oracle.hf_cls_states(model, ids, tt, lens), pw, pb)
    assert np.allclose(np.linalg.norm(got, axis=1), 1.0, atol=1e-5)
    assert np.all(np.linalg.norm(got - want, axis=1) <= 1e-3), np.linalg.norm(got - want, axis=1).max()
    assert np.allclose(got, want, rtol=1e-3, atol=3e-4), np.abs(got - want).max()
    # cosine between GPU and oracle embeddings
    assert np.all(np.sum(got * want, axis=1) > 1 - 1e-6)
    # raw (un-normalised) output too
    raw = engine.enc_embed(ids[:5], tt[:5], lens[:5], normalize=False)
    want_raw = ce_oracle.embed_from_cls(ce_oracle.hf_cls_states(model, ids[:5], tt[:5], lens[:5]), pw, pb, normalize=False)
    # un-normalised vectors: relative L2 error (fp16 GEMM operands, fp32 accumulation)
    assert np.all(np.linalg.norm(raw - want_raw, axis=1) <= 2e-3 * np.linalg.norm(want_raw, axis=1))


def test_embedder_class_surface_and_dense_chain(engine):
    import torch

    emb = B200Embedder(engine=engine, dimension=1024, seed=4, allow_random_init=True)
    assert emb.dimension == 1024
    corpus_texts = _texts(3000)
    texts = corpus_texts[:12]
    many = emb.embed_many_sync(texts)
    assert len(many) == 12 and all(len(v) == 1024 for v in many)
    one = emb.embed_sync(texts[3])
    assert one == many[3] and emb.stats["cache_hits"] >= 1
    assert asyncio.run(emb.embed_async_single(texts[4])) == many[4]
    assert asyncio.run(emb.warm_up()) is True
    # device chain: embed on the device -> dense top-k without a host round trip == host path
    corpus = np.asarray(emb.embed_arrays(corpus_texts), dtype=np.float32)
    engine.load_dense(corpus)
    ids, tt, lens = tokenize_for_embedding(texts, 128)
    q_dev = engine.enc_embed_dev(torch.from_numpy(ids).cuda(), torch.from_numpy(tt).cuda(), torch.from_numpy(lens).cuda())
    d_ids, d_sc, d_cnt = engine.dense_topk_dev(q_dev, 10)
    torch.cuda.synchronize()
    h_ids, h_sc, h_cnt = engine.dense_topk(np.asarray(many, np.float32), 10)
    assert np.array_equal(d_ids.cpu().numpy(), h_ids) and np.allclose(d_sc.cpu().numpy(), h_sc, rtol=1e-6, atol=1e-7)
    # every text's nearest neighbour in a corpus that contains it is itself
    assert list(h_ids[2:, 0]) == list(range(2, 12))
