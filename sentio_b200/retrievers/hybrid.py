"""HybridRetriever -- dense + sparse (+ plugin) retrieval fused on the GPU (K3).

Same constructor, attributes and observable behaviour as reference src/core/retrievers/hybrid.py:48-323:
every sub-retriever is queried with the caller's ``top_k``; cache-collection hits are prepended to the dense list;
fusion method ``rrf`` / ``weighted_rrf`` / ``comb_sum`` (unknown -> ``ValueError`` at query time); scorer plugins add
their scores to the merged documents; stable descending sort; truncate to ``top_k`` THEN drop ids without a document;
``dense_score`` / ``plugin_{i}_score`` / ``hybrid_score`` / ``score`` metadata side effects.

The arithmetic (rank / score fusion, accumulation order, tie-stable ranking) runs in ``sb_fuse`` (csrc/fuse.cu); this
class only marshals ids and keeps the reference's control flow and error policy (dense errors propagate, plugin errors
are swallowed).
"""
from __future__ import annotations

import logging
import os

import numpy as np

from ..document import Document
from ..engine import FUSION_METHODS, B200Engine
from .base import BaseRetriever, ScorerPlugin
from .dense import DenseRetriever
from .sparse import BM25Retriever

logger = logging.getLogger(__name__)


class HybridRetrieverPlugin:
    """External retriever plugin: ``retrieve(query, top_k) -> list[(doc_id, score)]`` (hybrid.py:27-45)."""

    def retrieve(self, query: str, top_k: int) -> list[tuple[str, float]]:
        raise NotImplementedError


class HybridRetriever(BaseRetriever):
    def __init__(self, dense_retriever: DenseRetriever, sparse_retriever: BaseRetriever | None = None,
                 corpus_docs: list[Document] | None = None, rrf_k: int = 60,
                 scorer_plugins: list[ScorerPlugin] | None = None,
                 retriever_plugins: list[HybridRetrieverPlugin] | None = None, use_pyserini: bool = False,
                 fusion_method: str = "rrf", dense_weight: float = 0.5, sparse_weight: float = 0.5,
                 engine: B200Engine | None = None, device: int = 0) -> None:
        self._dense = dense_retriever
        self._rrf_k = rrf_k
        self._scorer_plugins = scorer_plugins or []
        self._retriever_plugins = retriever_plugins or []
        self.fusion_method = fusion_method
        self.dense_weight = dense_weight
        self.sparse_weight = sparse_weight
        self._engine = engine
        self._device = device

        self._has_cache_collection = False
        self._cache_collection_name = os.getenv("CACHE_COLLECTION_NAME", "web_cache")
        try:
            client = getattr(self._dense, "_client", None)
            if client and hasattr(client, "collection_exists"):
                if client.collection_exists(collection_name=self._cache_collection_name):
                    self._has_cache_collection = True
        except Exception as exc:
            logger.warning("Failed to check cache collection: %s", exc)

        self._sparse_retriever = sparse_retriever
        if self._sparse_retriever is None and corpus_docs:
            # Pyserini (JVM) is out of scope on this path; the in-memory BM25 of the reference is the GPU BM25 here
            self._sparse_retriever = BM25Retriever(documents=corpus_docs, device=device)

    # ------------------------------------------------------------------ helpers
    def _fusion_engine(self) -> B200Engine:
        if self._engine is None:
            eng = getattr(self._sparse_retriever, "_engine", None)
            self._engine = eng if isinstance(eng, B200Engine) else B200Engine(self._device)
        return self._engine

    def _cache_hits(self, query: str, top_k: int) -> list[Document]:
        if not self._has_cache_collection:
            return []
        try:
            client = getattr(self._dense, "_client", None)
            embedder = getattr(self._dense, "_embedder", None)
            if not (client and embedder):
                return []
            cache_retriever = type(self._dense)(client=client, embedder=embedder,
                                                collection_name=self._cache_collection_name,
                                                vector_name=getattr(self._dense, "_vector_name", None))
            return cache_retriever.retrieve(query, top_k=top_k)
        except Exception as exc:
            logger.warning("Failed to retrieve from cache collection: %s", exc)
            return []

    # ------------------------------------------------------------------ public API
    def retrieve(self, query: str, top_k: int = 10) -> list[Document]:
        dense_hits = self._dense.retrieve(query, top_k=top_k)  # failures propagate, like the reference
        sparse_docs: list[Document] = []
        if self._sparse_retriever:
            sparse_docs = self._sparse_retriever.retrieve(query, top_k=top_k)
        return self._fuse(query, dense_hits, sparse_docs, top_k)

    def retrieve_batch(self, queries, top_k: int = 10) -> list[list[Document]]:
        """Many queries: the two retrieval stages run as ONE device batch each (``retrieve_batch`` of the dense and the
        sparse retriever -- the sparse one hands out per-query copies, see ``BM25Retriever.retrieve_batch``), the fusion
        of every query is the same code as ``retrieve``."""
        queries = list(queries)
        if not queries:
            return []
        dense_lists = self._dense.retrieve_batch(queries, top_k=top_k)
        if self._sparse_retriever:
            sparse_lists = self._sparse_retriever.retrieve_batch(queries, top_k=top_k)
        else:
            sparse_lists = [[] for _ in queries]
        return [self._fuse(q, d, s, top_k) for q, d, s in zip(queries, dense_lists, sparse_lists)]

    def _fuse(self, query: str, dense_hits: list[Document], sparse_docs: list[Document], top_k: int) -> list[Document]:
        """Everything of ``HybridRetriever.retrieve`` after the two retrieval calls (hybrid.py:146-300)."""
        all_dense = self._cache_hits(query, top_k) + dense_hits

        plugin_hits: list[tuple[str, float]] = []
        for plugin in self._retriever_plugins:
            try:
                plugin_hits.extend(plugin.retrieve(query, top_k))
            except Exception as exc:
                logger.warning("Retriever plugin failed: %s", exc)

        if self.fusion_method not in FUSION_METHODS:
            raise ValueError(f"Unknown fusion_method: {self.fusion_method}")

        # string ids -> dense integer codes (the kernel only needs id equality)
        code_of: dict[str, int] = {}

        def code(doc_id) -> int:
            c = code_of.get(doc_id)
            if c is None:
                c = code_of[doc_id] = len(code_of)
            return c

        d_ids = np.fromiter((code(d.id) for d in all_dense), dtype=np.int64, count=len(all_dense))
        d_sc = np.fromiter((float(d.metadata.get("score", 0.0)) for d in all_dense), dtype=np.float64,
                           count=len(all_dense))
        for doc, raw in zip(all_dense, d_sc):
            # rrf modes keep whatever object was stored; comb_sum stores the float (hybrid.py:228,232-234)
            doc.metadata["dense_score"] = float(raw) if self.fusion_method == "comb_sum" else doc.metadata.get("score", 0.0)
        s_ids = np.fromiter((code(d.id) for d in sparse_docs), dtype=np.int64, count=len(sparse_docs))
        s_sc = np.fromiter((float(d.metadata.get("bm25_score", 0.0)) for d in sparse_docs), dtype=np.float64,
                           count=len(sparse_docs))
        p_ids = np.fromiter((code(i) for i, _ in plugin_hits), dtype=np.int64, count=len(plugin_hits))
        p_sc = np.fromiter((float(s) for _, s in plugin_hits), dtype=np.float64, count=len(plugin_hits))

        # merged documents: dense (later duplicates replace the object, first position is kept), then sparse-only
        id_to_doc: dict[str, Document] = {}
        for doc in all_dense:
            id_to_doc[doc.id] = doc
        for doc in sparse_docs:
            if doc.id not in id_to_doc:
                id_to_doc[doc.id] = doc
        merged_docs = list(id_to_doc.values())

        extra_rows = []
        for plugin_idx, scorer in enumerate(self._scorer_plugins):
            try:
                plugin_scores = scorer.score(query, merged_docs)
                row = np.zeros(len(merged_docs), dtype=np.float64)
                for idx, (doc, score) in enumerate(zip(merged_docs, plugin_scores)):
                    doc.metadata[f"plugin_{plugin_idx}_score"] = float(score)
                    row[idx] = float(score)
                extra_rows.append(row)
            except Exception as exc:
                logger.warning("Scorer plugin %d failed: %s", plugin_idx, exc)
        extra = np.stack(extra_rows)[None, :, :] if extra_rows and merged_docs else None

        n_unique = len(code_of)
        if n_unique == 0:
            return []
        k = int(top_k)
        if k <= 0:
            return []

        def lst(ids, sc):
            if len(ids) == 0:
                return None
            return ids[None, :], sc[None, :], np.asarray([len(ids)], dtype=np.int32)

        ids, scores, src, counts = self._fusion_engine().fuse(
            self.fusion_method, float(self._rrf_k), float(self.dense_weight), float(self.sparse_weight), k,
            dense=lst(d_ids, d_sc), sparse=lst(s_ids, s_sc), plugin=lst(p_ids, p_sc), extra=extra)

        id_of_code = list(code_of.keys())
        results = []
        for j in range(int(counts[0])):
            doc = id_to_doc.get(id_of_code[int(ids[0, j])])
            if doc is None:  # plugin-only id: dropped after truncation (hybrid.py:291-298)
                continue
            score = float(scores[0, j])
            doc.metadata["hybrid_score"] = score
            doc.metadata["score"] = score
            results.append(doc)
        return results

    # ------------------------------------------------------------------ plugin management
    def add_scorer_plugin(self, scorer: ScorerPlugin) -> None:
        if scorer not in self._scorer_plugins:
            self._scorer_plugins.append(scorer)

    def add_retriever_plugin(self, retriever: HybridRetrieverPlugin) -> None:
        if retriever not in self._retriever_plugins:
            self._retriever_plugins.append(retriever)
