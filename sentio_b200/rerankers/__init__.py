"""Reranker package surface (mirrors reference src/core/rerankers/__init__.py:11-30)."""
from __future__ import annotations

from typing import Any

__all__ = ["get_reranker"]


def get_reranker(kind: str | None = None, **kwargs: Any):
    """``b200`` (default here) -> local GPU cross-encoder.  ``jina`` is the reference's remote reranker, not ours."""
    kind = (kind or "b200").lower()
    if kind in {"b200", "b200-minilm", "cross-encoder", "local"}:
        from .b200_reranker import B200Reranker

        return B200Reranker(**kwargs)
    raise ValueError(f"Unknown reranker kind: {kind}")
