"""Retriever / scorer interfaces (same surface as reference src/core/retrievers/base.py:13-42)."""
from __future__ import annotations

import abc
import asyncio
from typing import Protocol, runtime_checkable

from ..document import Document

__all__ = ["BaseRetriever", "ScorerPlugin"]


@runtime_checkable
class ScorerPlugin(Protocol):
    """Anything with ``score(query, docs) -> list[float]`` (one float per doc, same order)."""

    def score(self, query: str, docs: list[Document]) -> list[float]:
        ...


class BaseRetriever(abc.ABC):
    """``retrieve(query, top_k=10) -> list[Document]`` best first; ``retrieve_async`` runs it on the default executor."""

    @abc.abstractmethod
    def retrieve(self, query: str, top_k: int = 10) -> list[Document]:
        raise NotImplementedError

    async def retrieve_async(self, query: str, top_k: int = 10) -> list[Document]:
        loop = asyncio.get_running_loop()
        return await loop.run_in_executor(None, self.retrieve, query, top_k)
