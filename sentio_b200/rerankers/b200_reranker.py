"""B200Reranker -- a local BERT-style cross-encoder behind the reference's ``Reranker`` interface (K5).

Control flow follows ``JinaReranker.rerank`` (reference src/core/rerankers/jina_reranker.py:192-322): empty docs ->
``[]``; blank query -> default ranking; text falls back to ``metadata["content"]``; every document gets
``metadata["rerank_score"] = metadata["score"] = relevance in [0, 1]``; stable descending sort; ``[:top_k]``; NEVER
raises -- any failure degrades to the default ranking (input order, ``rerank_score = 1.0 - 0.1 * idx``, ``score``
untouched).  The remote call is replaced by ``sb_ce_score`` (csrc/cross_encoder.cu): sigmoid(logit) of a MiniLM-L6
shaped sequence classifier over ``[CLS] query [SEP] document [SEP]``.
"""
from __future__ import annotations

import logging
from typing import Any, Callable

import numpy as np

from ..cross_encoder import CrossEncoderWeights
from ..document import Document
from ..engine import B200Engine
from ..index import hash_tokenize_pairs
from .base import Reranker

logger = logging.getLogger(__name__)


class B200Reranker(Reranker):
    def __init__(self, weights: CrossEncoderWeights | None = None, engine: B200Engine | None = None, device: int = 0,
                 seq_len: int = 128, tokenizer: Callable | None = None, model_name: str = "b200-minilm-l6",
                 seed: int = 0, allow_random_init: bool = False):
        """``weights`` (and normally ``tokenizer``) are required: a reranker without trained weights reorders documents at
        random.  ``allow_random_init=True`` is for tests and benchmarks (BASELINE.json config 4 is a random-init model)."""
        self.model_name = model_name
        self.seq_len = int(seq_len)
        if weights is None:
            if not allow_random_init:
                raise ValueError("B200Reranker needs weights=CrossEncoderWeights(...) (e.g. from_hf_state_dict); pass "
                                 "allow_random_init=True only for tests / benchmarks")
            logger.warning("B200Reranker: RANDOM-INIT MiniLM-L6 weights (seed %d) -- scores are meaningless", seed)
            weights = CrossEncoderWeights.random_minilm_l6(seed=seed)
        if tokenizer is None:
            logger.warning("B200Reranker: no tokenizer given, using the crc32 hashing tokeniser (benchmark / test only)")
        self._tokenize = tokenizer or hash_tokenize_pairs
        self._engine = engine or B200Engine(device)
        self.weights = weights
        self._engine.ce_load(self.weights.blob(), self.weights.config)

    # ------------------------------------------------------------------ scoring
    def score_pairs(self, query: str, texts: list[str]) -> np.ndarray:
        ids, tt, lens = self._tokenize(query, texts, self.seq_len)
        _, sig = self._engine.ce_score(ids, tt, lens)
        return sig

    def rerank(self, query: str, docs: list[Document], top_k: int = 5, **kwargs: Any) -> list[Document]:
        if not docs:
            return []
        if not query or query.strip() == "":
            return self._default_ranking(docs, top_k)
        try:
            texts = []
            for doc in docs:
                text = doc.text
                if not text and doc.metadata and "content" in doc.metadata:
                    text = doc.metadata["content"]
                texts.append(text)
            scores = self.score_pairs(query, texts)
            for doc, score in zip(docs, scores):
                if not doc.text and doc.metadata and "content" in doc.metadata:
                    doc.text = doc.metadata["content"]
                doc.metadata["rerank_score"] = float(score)
                doc.metadata["score"] = float(score)
            ranked = sorted(docs, key=lambda d: d.metadata.get("rerank_score", 0.0), reverse=True)
            return ranked[:top_k]
        except Exception as exc:
            logger.error("B200 reranker failed, falling back to original order: %s", exc)
            return self._default_ranking(docs, top_k)

    def rerank_batch(self, queries, docs_per_query, top_k: int = 5, **kwargs: Any) -> list[list[Document]]:
        """All (query, document) pairs of all jobs in ONE cross-encoder forward; per job the same semantics as ``rerank``
        (blank queries / empty lists / failures degrade per job, never raise)."""
        if len(queries) != len(docs_per_query):
            raise ValueError("queries and docs_per_query must have the same length")
        out: list[list[Document] | None] = [None] * len(queries)
        jobs = []
        for i, (q, docs) in enumerate(zip(queries, docs_per_query)):
            if not docs:
                out[i] = []
            elif not q or q.strip() == "":
                out[i] = self._default_ranking(docs, top_k)
            else:
                jobs.append(i)
        if jobs:
            try:
                parts = []
                for i in jobs:
                    texts = [d.text if d.text or not (d.metadata and "content" in d.metadata) else d.metadata["content"]
                             for d in docs_per_query[i]]
                    parts.append(self._tokenize(queries[i], texts, self.seq_len))
                ids = np.concatenate([p[0] for p in parts])
                tt = np.concatenate([p[1] for p in parts])
                lens = np.concatenate([p[2] for p in parts])
                _, sig = self._engine.ce_score(ids, tt, lens)
                at = 0
                for i in jobs:
                    docs = docs_per_query[i]
                    for doc, score in zip(docs, sig[at:at + len(docs)]):
                        if not doc.text and doc.metadata and "content" in doc.metadata:
                            doc.text = doc.metadata["content"]
                        doc.metadata["rerank_score"] = float(score)
                        doc.metadata["score"] = float(score)
                    at += len(docs)
                    out[i] = sorted(docs, key=lambda d: d.metadata.get("rerank_score", 0.0), reverse=True)[:top_k]
            except Exception as exc:
                logger.error("B200 reranker batch failed, falling back to original order: %s", exc)
                for i in jobs:
                    if out[i] is None:
                        out[i] = self._default_ranking(docs_per_query[i], top_k)
        return out  # type: ignore[return-value]

    def _default_ranking(self, docs: list[Document], top_k: int) -> list[Document]:
        result = docs[:top_k]
        for idx, doc in enumerate(result):
            if not doc.text and doc.metadata and "content" in doc.metadata:
                doc.text = doc.metadata["content"]
            doc.metadata["rerank_score"] = 1.0 - (idx * 0.1)
        return result

    def get_health_status(self) -> dict[str, Any]:
        return {"service": "b200_reranker", "model": self.model_name, "is_healthy": True,
                "device": self._engine.device, "seq_len": self.seq_len}
