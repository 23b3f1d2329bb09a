"""Retriever package surface: the two registration points of the reference (``get_retriever`` / ``get_scorer``,
src/core/retrievers/__init__.py:17-79) backed by the GPU classes of this package.

Kinds are resolved through small registries (kind -> "module:Class"), imported lazily so that ``import
sentio_b200.retrievers`` never touches CUDA.
"""
from __future__ import annotations

from importlib import import_module
from typing import Any

from .base import BaseRetriever, ScorerPlugin

__all__ = ["BaseRetriever", "ScorerPlugin", "get_retriever", "get_scorer"]

_RETRIEVERS = {
    "dense": "dense:DenseRetriever",
    "hybrid": "hybrid:HybridRetriever",
    "bm25": "sparse:BM25Retriever",
    "sparse": "sparse:BM25Retriever",
    # same constructor contract as the reference's Lucene-backed class: RuntimeError without an index directory / corpus
    "pyserini": "sparse:PyseriniBM25Retriever",
    "lucene": "sparse:PyseriniBM25Retriever",
}

_SCORERS = {
    "keyword": "KeywordMatchScorer",
    "recency": "RecencyScorer",
    "time": "RecencyScorer",
    "semantic": "SemanticSimilarityScorer",
    "similarity": "SemanticSimilarityScorer",
    "mmr": "MMRScorer",
}


def _resolve(spec: str):
    module, _, name = spec.partition(":")
    return getattr(import_module(f"{__name__}.{module}"), name)


def get_retriever(kind: str, **kwargs: Any):
    """Retriever instance for ``kind`` (case-insensitive); keyword arguments go to the constructor.

    ``hybrid`` follows the reference's rules: a ``top_k`` keyword is dropped (it is not a constructor parameter) and
    ``dense_retriever`` must be passed explicitly."""
    key = kind.lower()
    spec = _RETRIEVERS.get(key)
    if spec is None:
        raise ValueError(f"Unknown retriever kind: {kind}")
    if key == "hybrid":
        kwargs.pop("top_k", None)
        if "dense_retriever" not in kwargs:
            raise ValueError("For hybrid retriever, 'dense_retriever' must be provided explicitly")
    return _resolve(spec)(**kwargs)


def get_scorer(kind: str, **kwargs: Any) -> ScorerPlugin:
    name = _SCORERS.get(kind.lower())
    if name is None:
        raise ValueError(f"Unknown scorer kind: {kind}")
    return getattr(import_module(f"{__name__}.scorers"), name)(**kwargs)
