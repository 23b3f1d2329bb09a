"""Document selector -- the node that follows the reranker.

Mirror of ``create_document_selector_node`` (reference src/core/graph/nodes.py:231-372): take the reranked (else the
retrieved) documents, stable-sort by ``metadata["score"]`` descending, drop repeated ids, walk the first ``top_k`` of
them, fall back to ``metadata["content"]`` for empty texts, skip blank documents, and keep adding documents while the
running ``len(text) // 4`` token estimate stays within ``max_tokens`` (the first document that does not fit ends the
walk).  Same state keys (``selected_documents``, ``metadata.selected_count`` / ``selected_tokens`` /
``selector_error``), same ``metadata.user_top_k`` override, same exception fallback.

``select_documents`` is the single-request host form (<= top_k strings: nothing to accelerate); the batched device form
over candidate-id arrays is ``B200Engine.select_dev`` (``sb_select_dev``), which folds the same walk behind the rerank
kernels so a batched pipeline returns final context documents without a host round trip.
"""
from __future__ import annotations

import logging
from typing import Any, Callable

from .document import Document

logger = logging.getLogger(__name__)


def _usable_text(doc: Document):
    text = doc.text
    if not text and doc.metadata and "content" in doc.metadata:
        text = doc.metadata["content"]
    return text


def select_documents(candidate_docs: list[Document], top_k: int = 3, max_tokens: int = 2000):
    """-> (selected fresh ``Document`` copies, total estimated tokens); nodes.py:272-337."""
    sorted_candidates = sorted(candidate_docs, key=lambda d: float(d.metadata.get("score", 0.0) or 0.0), reverse=True)
    seen_ids: set[str] = set()
    unique_candidates: list[Document] = []
    for doc in sorted_candidates:
        if doc.id and doc.id in seen_ids:
            continue
        if doc.id:
            seen_ids.add(doc.id)
        unique_candidates.append(doc)
    selected: list[Document] = []
    total_tokens = 0
    for doc in unique_candidates[:top_k]:
        text = _usable_text(doc)
        if not text or not str(text).strip():
            continue
        doc_tokens = len(text) // 4
        if total_tokens + doc_tokens <= max_tokens:
            selected.append(Document(id=doc.id, text=text, metadata=doc.metadata.copy() if doc.metadata else {}))
            total_tokens += doc_tokens
        else:
            break
    return selected, total_tokens


def selector_chars(doc: Document) -> int:
    """Per-document input of the device selector: characters of the usable text, 0 for a blank document."""
    text = _usable_text(doc)
    if not text or not str(text).strip():
        return 0
    return len(text)


def create_document_selector_node(top_k: int = 3, max_tokens: int = 2000) -> Callable[[dict], dict]:
    def select_documents_node(state: dict[str, Any]) -> dict[str, Any]:
        candidate_docs = state["reranked_documents"] or state["retrieved_documents"]
        if not candidate_docs:
            logger.warning("No documents to select")
            return state
        try:
            user = state.get("metadata", {}).get("user_top_k", top_k)
            effective_top_k = int(user) if isinstance(user, (int, float)) else top_k
            selected, total_tokens = select_documents(candidate_docs, effective_top_k, max_tokens)
            state["selected_documents"].extend(selected)
            state["metadata"]["selected_count"] = len(selected)
            state["metadata"]["selected_tokens"] = total_tokens
        except Exception as exc:
            logger.error("Error selecting documents: %s", exc)
            state["metadata"]["selector_error"] = str(exc)
            fallback = []
            for doc in candidate_docs[:min(top_k, len(candidate_docs))]:
                fallback.append(Document(id=doc.id, text=_usable_text(doc),
                                         metadata=doc.metadata.copy() if doc.metadata else {}))
            state["selected_documents"].extend(fallback)
        return state

    return select_documents_node
