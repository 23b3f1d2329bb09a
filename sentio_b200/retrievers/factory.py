"""Environment-driven construction (same knobs as reference src/core/retrievers/factory.py:21-196).

``RETRIEVAL_STRATEGY`` (dense | bm25 | pyserini | hybrid), ``BM25_INDEX_DIR`` / ``BM25_K1`` / ``BM25_B``, ``RRF_K`` (default 20 here, like the reference factory),
``FUSION_METHOD``, ``DENSE_WEIGHT``, ``SPARSE_WEIGHT``, ``BM25_VARIANT``, ``COLLECTION_NAME``, ``TEXT_VECTOR_NAME``.
When no scorer plugins are passed the reference's default trio is installed (semantic 0.8, keyword 0.2, MMR 0.5/0.5).
The BM25 corpus is pulled from the store with ``scroll`` exactly like the reference does against Qdrant.
"""
from __future__ import annotations

import logging
import os

from ..document import Document
from .base import BaseRetriever, ScorerPlugin
from .dense import DenseRetriever
from .hybrid import HybridRetriever
from .sparse import BM25Retriever, PyseriniBM25Retriever

logger = logging.getLogger(__name__)


def _scroll_corpus(client, collection_name: str) -> list[Document]:
    docs: list[Document] = []
    offset = None
    raw = getattr(client, "_client", client)
    while True:
        points, offset = raw.scroll(collection_name=collection_name, with_payload=True, with_vectors=False, limit=100,
                                    offset=offset)
        if not points:
            break
        for point in points:
            payload = point.payload
            if payload:
                text = (payload.get("text") or payload.get("content") or payload.get("document")
                        or payload.get("page_content") or "")
                docs.append(Document(id=str(point.id), text=text, metadata=payload.get("metadata", {})))
        if offset is None:
            break
    return docs


def create_retriever_from_env(client, embedder, collection_name: str | None = None,
                              corpus_docs: list[Document] | None = None,
                              scorer_plugins: list[ScorerPlugin] | None = None) -> BaseRetriever:
    strategy = os.getenv("RETRIEVAL_STRATEGY", "dense").lower()
    rrf_k = int(os.getenv("RRF_K", "20"))
    bm25_variant = os.getenv("BM25_VARIANT", "okapi").lower()
    fusion_method = os.getenv("FUSION_METHOD", "rrf").lower()
    try:
        dense_weight = float(os.getenv("DENSE_WEIGHT", "0.5"))
        sparse_weight = float(os.getenv("SPARSE_WEIGHT", "0.5"))
    except ValueError:
        dense_weight, sparse_weight = 0.5, 0.5
    if collection_name is None:
        collection_name = os.getenv("COLLECTION_NAME", "Sentio_docs")
    vector_name = os.getenv("TEXT_VECTOR_NAME", "text-dense")

    dense = DenseRetriever(client=client, embedder=embedder, collection_name=collection_name, vector_name=vector_name)

    if scorer_plugins is None:
        from .scorers import KeywordMatchScorer, MMRScorer, SemanticSimilarityScorer

        scorer_plugins = [SemanticSimilarityScorer(embedder=embedder, weight=0.8), KeywordMatchScorer(weight=0.2),
                          MMRScorer(embedder=embedder, lambda_=0.5, weight=0.5)]

    if corpus_docs is None and strategy in ("hybrid", "bm25", "pyserini"):
        try:
            corpus_docs = _scroll_corpus(client, collection_name)
        except Exception as exc:
            logger.error("Failed to load documents from the vector store: %s", exc)
            corpus_docs = []

    if strategy == "dense":
        return dense
    if strategy == "bm25":
        return BM25Retriever(documents=corpus_docs or [], variant=bm25_variant,
                             cache_dir=os.getenv("SPARSE_CACHE_DIR", ".sparse_cache"))
    if strategy == "pyserini":  # factory.py:150-163: try the Lucene retriever, fall back to BM25 when it cannot start
        try:
            return PyseriniBM25Retriever(index_dir=os.getenv("BM25_INDEX_DIR", "indexes/lucene-index"),
                                         k1=float(os.getenv("BM25_K1", "0.9")), b=float(os.getenv("BM25_B", "0.4")))
        except RuntimeError as exc:
            logger.error("Failed to initialize Pyserini, falling back to BM25: %s", exc)
            return BM25Retriever(documents=corpus_docs or [], variant=bm25_variant)
    if strategy == "hybrid":
        sparse = None
        index_dir = os.getenv("BM25_INDEX_DIR", "indexes/lucene-index")
        if os.path.isdir(index_dir):  # factory.py:168-176: prefer Pyserini when its index exists
            try:
                sparse = PyseriniBM25Retriever(index_dir=index_dir)
            except RuntimeError as exc:
                logger.error("Failed to initialize Pyserini: %s", exc)
        if sparse is None:
            sparse = BM25Retriever(documents=corpus_docs, variant=bm25_variant) if corpus_docs else None
        if sparse is None:
            logger.warning("Hybrid search selected, but no sparse retriever available.")
        return HybridRetriever(dense_retriever=dense, corpus_docs=corpus_docs, rrf_k=rrf_k,
                               scorer_plugins=scorer_plugins, sparse_retriever=sparse, fusion_method=fusion_method,
                               dense_weight=dense_weight, sparse_weight=sparse_weight)
    raise ValueError(f"Unknown retrieval strategy: {strategy}")
