"""Retriever package surface (mirrors reference src/core/retrievers/__init__.py:17-79: ``get_retriever`` / ``get_scorer``)."""
from __future__ import annotations

from typing import Any

from .base import BaseRetriever, ScorerPlugin

__all__ = ["BaseRetriever", "ScorerPlugin", "get_retriever", "get_scorer"]


def get_retriever(kind: str, **kwargs: Any):
    """``dense`` | ``hybrid`` | ``bm25``/``sparse`` | ``pyserini``/``lucene`` -> GPU-backed retriever; same argument rules as the reference."""
    kind = kind.lower()
    if kind == "dense":
        from .dense import DenseRetriever

        return DenseRetriever(**kwargs)
    if kind == "hybrid":
        from .hybrid import HybridRetriever

        kwargs.pop("top_k", None)
        if "dense_retriever" not in kwargs:
            raise ValueError("For hybrid retriever, 'dense_retriever' must be provided explicitly")
        return HybridRetriever(**kwargs)
    if kind in ("bm25", "sparse"):
        from .sparse import BM25Retriever

        return BM25Retriever(**kwargs)
    if kind in ("pyserini", "lucene"):
        from .sparse import PyseriniBM25Retriever

        return PyseriniBM25Retriever(**kwargs)  # RuntimeError without an index dir / corpus, like the reference
    raise ValueError(f"Unknown retriever kind: {kind}")


def get_scorer(kind: str, **kwargs: Any) -> ScorerPlugin:
    from . import scorers

    kind = kind.lower()
    if kind == "keyword":
        return scorers.KeywordMatchScorer(**kwargs)
    if kind in ("recency", "time"):
        return scorers.RecencyScorer(**kwargs)
    if kind in ("semantic", "similarity"):
        return scorers.SemanticSimilarityScorer(**kwargs)
    if kind == "mmr":
        return scorers.MMRScorer(**kwargs)
    raise ValueError(f"Unknown scorer kind: {kind}")
