"""Exact-cosine dense top-k oracle (TEST INFRASTRUCTURE).

Restates what the reference obtains from Qdrant at src/core/retrievers/dense.py:46-64 for a collection created with
distance="Cosine" (src/core/vector_store/qdrant_store.py:51-52): best-first cosine similarity, evaluated here EXACTLY
(fp64) -- Qdrant itself is approximate (HNSW) and fp32, so exact cosine is the parity target.

Storage semantics shared with the product (DESIGN.md "K1"): fp32 input vectors are L2-normalised and rounded to fp16
("normalise at upsert"); fp16 input is stored verbatim; the cosine is computed on the STORED values.
"""
from __future__ import annotations

import numpy as np


def stored_rows(vecs: np.ndarray) -> np.ndarray:
    """What the index holds for these input vectors (fp16)."""
    v = np.asarray(vecs)
    if v.dtype == np.float16:
        return v
    v64 = v.astype(np.float32).astype(np.float64)
    nrm = np.sqrt((v64 * v64).sum(axis=1, keepdims=True))
    nrm[nrm == 0.0] = 1.0
    return (v64 / nrm).astype(np.float16)


def cosine_scores(rows16: np.ndarray, q: np.ndarray) -> np.ndarray:
    """fp64 cos(q, row) for every stored row; 0 where a norm is 0."""
    x = rows16.astype(np.float64)
    q64 = np.asarray(q, dtype=np.float32).astype(np.float64)
    dot = x @ q64
    den = np.sqrt((x * x).sum(axis=1)) * np.sqrt((q64 * q64).sum())
    out = np.zeros(len(x))
    np.divide(dot, den, out=out, where=den > 0)
    return out


def topk(scores: np.ndarray, k: int):
    """(score desc, index asc) -- the deterministic total order the product implements."""
    order = np.lexsort((np.arange(len(scores)), -scores))[:k]
    return order, scores[order]


def dense_topk(rows16: np.ndarray, q: np.ndarray, k: int, chunk: int = 131072):
    """Chunked exact top-k for big corpora: returns (indices, scores)."""
    n = len(rows16)
    if n <= chunk:
        return topk(cosine_scores(rows16, q), k)
    best_i, best_s = np.zeros(0, np.int64), np.zeros(0)
    for lo in range(0, n, chunk):
        s = cosine_scores(rows16[lo:lo + chunk], q)
        i, v = topk(s, k)
        best_i = np.concatenate([best_i, i + lo])
        best_s = np.concatenate([best_s, v])
        o = np.lexsort((best_i, -best_s))[:k]
        best_i, best_s = best_i[o], best_s[o]
    return best_i, best_s


def dense_topk_multi(rows16: np.ndarray, Q: np.ndarray, k: int, chunk: int = 131072):
    """Exact top-k for several queries at once (each corpus chunk is widened to fp64 once): list of (indices, scores)."""
    Q64 = np.asarray(Q, dtype=np.float32).astype(np.float64)
    qn = np.sqrt((Q64 * Q64).sum(axis=1))
    best = [(np.zeros(0, np.int64), np.zeros(0)) for _ in range(len(Q64))]
    for lo in range(0, len(rows16), chunk):
        x = rows16[lo:lo + chunk].astype(np.float64)
        xn = np.sqrt((x * x).sum(axis=1))
        for b in range(len(Q64)):
            den = xn * qn[b]
            s = np.zeros(len(x))
            np.divide(x @ Q64[b], den, out=s, where=den > 0)
            i, v = topk(s, k)
            bi = np.concatenate([best[b][0], i + lo])
            bs = np.concatenate([best[b][1], v])
            o = np.lexsort((bi, -bs))[:k]
            best[b] = (bi[o], bs[o])
    return best


def fast_topk_f32(x32: np.ndarray, q: np.ndarray, k: int):
    """The reference-CPU-arm implementation timed by bench.py: fp32 `X @ q` (BLAS, all cores) on pre-normalised rows,
    then the best-first top-k the way the reference code base does it (full ``np.argsort``, cf. sparse.py:180)."""
    s = x32 @ q
    idx = np.argsort(-s)[:k]
    return idx, s[idx]
