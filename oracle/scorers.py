"""Scorer oracles (TEST INFRASTRUCTURE): NumPy restatements of SemanticSimilarityScorer.score and MMRScorer.score,
reference src/core/retrievers/scorers.py:152-191 and :222-273, taking the embeddings directly."""
from __future__ import annotations

import numpy as np


def _cos(a, b) -> float:
    denom = np.linalg.norm(a) * np.linalg.norm(b)
    return float(np.dot(a, b) / denom) if denom else 0.0


def semantic(query_emb, doc_embs, weight: float):
    qn = np.linalg.norm(query_emb)
    out = []
    for d in doc_embs:
        dn = np.linalg.norm(d)
        out.append(float(np.dot(query_emb, d) / (qn * dn) * weight) if (qn > 0 and dn > 0) else 0.0)
    return out


def mmr(query_emb, doc_embs, lambda_: float, weight: float):
    n = len(doc_embs)
    if n == 0:
        return []
    rel = [_cos(query_emb, d) for d in doc_embs]
    sim = {}

    def cos_ij(i, j):
        key = (i, j) if i < j else (j, i)
        if key not in sim:
            sim[key] = _cos(doc_embs[i], doc_embs[j])
        return sim[key]

    scores = [0.0] * n
    selected: list[int] = []
    for _ in range(n):
        best_idx, best = None, -1.0
        for idx in range(n):
            if idx in selected:
                continue
            redundancy = 0.0
            for s in selected:
                redundancy = max(redundancy, cos_ij(idx, s))
            val = lambda_ * rel[idx] - (1 - lambda_) * redundancy
            if val > best:
                best, best_idx = val, idx
        if best_idx is None:
            break
        selected.append(best_idx)
        scores[best_idx] = best * weight
    for idx in range(n):
        if scores[idx] == 0.0:
            scores[idx] = rel[idx] * weight * lambda_
    return [max(0.0, s) for s in scores]
