"""DenseRetriever -- exact cosine top-k over an HBM-resident fp16 corpus (K1).

Same constructor and result contract as reference src/core/retrievers/dense.py:21-119
(``DenseRetriever(client, embedder, collection_name, vector_name=None)``; results are ``Document(id=str(point.id),
text=payload text|content|metadata.content, metadata={"score": point.score, **payload})``), but ``client`` is a
``B200VectorStore`` whose ``search`` is a brute-force scan on the GPU instead of a Qdrant round trip.
"""
from __future__ import annotations

import inspect
import logging

from ..document import Document
from .base import BaseRetriever

logger = logging.getLogger(__name__)

_DEFAULT_VECTOR_NAME = "text-dense"


def payload_text(payload: dict) -> str:
    """Text resolution order of the reference: payload["text"], payload["content"], payload["metadata"]["content"]."""
    text = payload.get("text") or payload.get("content") or ""
    if not text:
        nested = payload.get("metadata")
        if isinstance(nested, dict):
            text = nested.get("content", "")
    return text


class DenseRetriever(BaseRetriever):
    def __init__(self, client, embedder, collection_name: str, vector_name: str | None = None,
                 document_cls: type = Document) -> None:
        # a store wrapper exposing ``_client`` is unwrapped like the reference does for QdrantStore (dense.py:31-33)
        self._client = getattr(client, "_client", client)
        self._embedder = embedder
        self._collection = collection_name
        self._vector_name = vector_name or _DEFAULT_VECTOR_NAME
        self._document_cls = document_cls

    def retrieve(self, query: str, top_k: int = 10) -> list[Document]:
        query_vec = self._embedder.embed_sync(query)
        kwargs = dict(collection_name=self._collection, query_vector=query_vec, limit=top_k, with_payload=True,
                      with_vectors=False)
        try:
            if "vector_name" in inspect.signature(self._client.search).parameters:
                kwargs["vector_name"] = self._vector_name
            points = self._client.search(**kwargs)
        except TypeError as exc:
            if "vector_name" not in str(exc):
                raise
            kwargs.pop("vector_name", None)
            points = self._client.search(**kwargs)

        docs = self._documents(points)
        logger.debug("DenseRetriever: %d documents for top_k=%d", len(docs), top_k)
        return docs

    def _documents(self, points) -> list[Document]:
        docs = []
        for point in points:
            payload = point.payload or {}
            text = payload_text(payload)
            metadata = {"score": point.score, **payload}
            if text and "content" not in metadata:
                metadata["content"] = text
            docs.append(self._document_cls(id=str(point.id), text=text, metadata=metadata))
        return docs

    def retrieve_batch(self, queries, top_k: int = 10) -> list[list[Document]]:
        """Many queries, one embedding call and ONE device scan when the store offers ``search_batch`` (B200VectorStore);
        any other client falls back to the per-query loop of the base class."""
        queries = list(queries)
        if not queries:
            return []
        search_batch = getattr(self._client, "search_batch", None)
        embed_many = getattr(self._embedder, "embed_many_sync", None)
        if search_batch is None or embed_many is None:
            return super().retrieve_batch(queries, top_k=top_k)
        vectors = embed_many(queries)
        hits = search_batch(collection_name=self._collection, query_vectors=vectors, limit=top_k, with_payload=True)
        return [self._documents(points) for points in hits]
