"""``Document`` -- the only data type that crosses the retriever / reranker boundary
(mirrors reference src/core/models/document.py:9-20: ``text``, ``metadata``, ``id`` with a uuid4 default).

Every class in this package is duck-typed on ``.text / .metadata / .id`` and accepts a ``document_cls`` so that, inside
the reference application, results can be built from the reference's own ``src.core.models.document.Document``.
"""
from __future__ import annotations

import uuid
from dataclasses import dataclass, field
from typing import Any, Dict

__all__ = ["Document"]


def _new_id() -> str:
    return str(uuid.uuid4())


@dataclass
class Document:
    text: str
    metadata: Dict[str, Any] = field(default_factory=dict)
    id: str = field(default_factory=_new_id)
