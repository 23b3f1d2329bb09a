"""Fusion oracle (TEST INFRASTRUCTURE): a plain-Python restatement of the fusion block of
HybridRetriever.retrieve, reference src/core/retrievers/hybrid.py:204-298, over (id, score) lists.

fuse(...) returns [(id, fused_score, has_doc)] best first, truncated to top_k; ``has_doc`` is False for ids that only a
retriever plugin produced (the reference drops those AFTER truncation, hybrid.py:291-298).
"""
from __future__ import annotations


def _normalise(values: dict) -> dict:
    # hybrid.py:211-220: min-max over the dict's values; all equal -> 1.0 for everybody
    if not values:
        return {}
    lo, hi = min(values.values()), max(values.values())
    if hi <= lo:
        return {k: 1.0 for k in values}
    scale = hi - lo
    return {k: (v - lo) / scale for k, v in values.items()}


def fuse(method, rrf_k, dense_weight, sparse_weight, dense, sparse, plugin, top_k, extras=None):
    """dense / sparse / plugin: lists of (id, raw_score) in retrieval order (cache hits already prepended to dense).

    extras: optional list (one entry per scorer plugin) of score lists aligned with the merged-document order
    (unique dense ids in first-occurrence order, then sparse-only ids; hybrid.py:262-271)."""
    fused: dict = {}

    def add(key, value):
        fused[key] = fused.get(key, 0.0) + value  # defaultdict(float) += : first touch inserts 0.0 then adds

    if method in ("rrf", "weighted_rrf"):
        for rank, (doc_id, _s) in enumerate(dense):  # hybrid.py:224-226
            w = 1.0 if method == "rrf" else float(dense_weight)
            add(doc_id, w * (1.0 / (rrf_k + rank)))
        for rank, (doc_id, _s) in enumerate(sparse):  # hybrid.py:242-244
            w = 1.0 if method == "rrf" else float(sparse_weight)
            add(doc_id, w * (1.0 / (rrf_k + rank)))
        for rank, (doc_id, _s) in enumerate(plugin):  # hybrid.py:254-255
            add(doc_id, 1.0 / (rrf_k + rank))
    elif method == "comb_sum":
        for lst, weight in ((dense, float(dense_weight)), (sparse, float(sparse_weight)), (plugin, 0.2)):
            raw = {}
            for doc_id, s in lst:
                raw[doc_id] = float(s)  # later duplicates overwrite, position of the first insertion is kept
            for doc_id, ns in _normalise(raw).items():
                add(doc_id, weight * ns)
    else:
        raise ValueError(f"Unknown fusion_method: {method}")

    merged = []
    seen = set()
    for doc_id, _ in dense:
        if doc_id not in seen:
            seen.add(doc_id)
            merged.append(doc_id)
    for doc_id, _ in sparse:
        if doc_id not in seen:
            seen.add(doc_id)
            merged.append(doc_id)
    for scores in extras or []:  # hybrid.py:275-285
        for doc_id, s in zip(merged, scores):
            add(doc_id, float(s))

    ranked = sorted(fused.items(), key=lambda kv: kv[1], reverse=True)[:top_k]  # stable
    return [(doc_id, score, doc_id in seen) for doc_id, score in ranked]
