"""Reranker interface (same surface as reference src/core/rerankers/base.py:14-132)."""
from __future__ import annotations

import abc
import asyncio
import functools
from typing import Any, Protocol

from ..document import Document

__all__ = ["Reranker", "RerankerProtocol", "RerankingResult"]


class RerankerProtocol(Protocol):
    def rerank(self, query: str, docs: list[Document], top_k: int = 5, **kwargs: Any) -> list[Document]:
        ...

    async def rerank_async(self, query: str, docs: list[Document], top_k: int = 5, **kwargs: Any) -> list[Document]:
        ...


class RerankingResult:
    """Reranked documents plus the originals and free-form metadata (sequence-like over ``documents``)."""

    def __init__(self, documents: list[Document], original_documents: list[Document] | None = None,
                 metadata: dict[str, Any] | None = None):
        self.documents = documents
        self.original_documents = original_documents or []
        self.metadata = metadata or {}

    @property
    def top_document(self) -> Document | None:
        return self.documents[0] if self.documents else None

    def __len__(self) -> int:
        return len(self.documents)

    def __getitem__(self, idx: int) -> Document:
        return self.documents[idx]

    def __iter__(self):
        return iter(self.documents)


class Reranker(abc.ABC):
    @abc.abstractmethod
    def rerank(self, query: str, docs: list[Document], top_k: int = 5, **kwargs: Any) -> list[Document]:
        raise NotImplementedError

    async def rerank_async(self, query: str, docs: list[Document], top_k: int = 5, **kwargs: Any) -> list[Document]:
        loop = asyncio.get_running_loop()
        return await loop.run_in_executor(None, functools.partial(self.rerank, query, docs, top_k, **kwargs))
