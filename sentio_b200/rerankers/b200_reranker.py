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
                 seed: int = 0):
        self.model_name = model_name
        self.seq_len = int(seq_len)
        self._tokenize = tokenizer or hash_tokenize_pairs
        self._engine = engine or B200Engine(device)
        self.weights = weights or CrossEncoderWeights.random_minilm_l6(seed=seed)
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
