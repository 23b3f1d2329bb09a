"""Seeded synthetic corpus / query generators shared by bench.py, the tests and the CPU baseline.

* dense corpus: ``default_rng(1234).standard_normal((N, D), float32)`` row-normalised then ROUNDED TO fp16 -- the rounded
  matrix *is* the corpus (so nothing is lost on upload and the fp64 oracle sees exactly what the GPU stores).
* texts: vocabulary of 50 000 tokens ``w0..w49999`` with p(rank r) ~ r^-1.07, doc length max(8, Poisson(80)) (seed 1235);
  generated as integer token streams (``doc_offsets``/``flat_tokens``); strings are materialised only on demand.
* queries: unit vectors (seed 4321) and 6-token texts from the same Zipf (seed 4322).
"""
from __future__ import annotations

import numpy as np

VOCAB = 50_000
ZIPF_S = 1.07


def dense_corpus(n: int, d: int, seed: int = 1234, chunk: int = 65536) -> np.ndarray:
    """fp16 [n, d]; generated in chunks with a generator that is advanced sequentially (prefix-stable in n)."""
    rng = np.random.default_rng(seed)
    out = np.empty((n, d), dtype=np.float16)
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        x = rng.standard_normal((hi - lo, d), dtype=np.float32)
        x /= np.linalg.norm(x, axis=1, keepdims=True)
        out[lo:hi] = x.astype(np.float16)
    return out


def dense_corpus_range(lo: int, hi: int, d: int, seed: int = 1234, chunk: int = 65536) -> np.ndarray:
    """Rows [lo, hi) of a corpus whose 65536-row chunks are seeded independently (SeedSequence([seed, chunk_index])):
    any shard can be generated without generating the rows before it (used for the 8 M-doc sharded runs).
    NOTE: a different stream than ``dense_corpus`` (which advances one generator sequentially)."""
    out = np.empty((hi - lo, d), dtype=np.float16)
    c = lo // chunk
    while c * chunk < hi:
        rng = np.random.default_rng(np.random.SeedSequence([seed, c]))
        x = rng.standard_normal((chunk, d), dtype=np.float32)
        x /= np.linalg.norm(x, axis=1, keepdims=True)
        a, b = max(lo, c * chunk), min(hi, (c + 1) * chunk)
        out[a - lo:b - lo] = x[a - c * chunk:b - c * chunk].astype(np.float16)
        c += 1
    return out


def query_vectors(b: int, d: int, seed: int = 4321) -> np.ndarray:
    rng = np.random.default_rng(seed)
    q = rng.standard_normal((b, d), dtype=np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    return q.astype(np.float32)


def _zipf_cdf(vocab: int = VOCAB, s: float = ZIPF_S) -> np.ndarray:
    p = np.arange(1, vocab + 1, dtype=np.float64) ** (-s)
    p /= p.sum()
    return np.cumsum(p)


def text_corpus_tokens(n: int, seed: int = 1235, vocab: int = VOCAB, mean_len: int = 80, min_len: int = 8):
    """-> (flat_tokens int32 [total], doc_offsets int64 [n+1]) ; token id r = Zipf rank r (0 = most frequent)."""
    rng = np.random.default_rng(seed)
    lens = np.maximum(min_len, rng.poisson(mean_len, size=n)).astype(np.int64)
    off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    cdf = _zipf_cdf(vocab)
    u = rng.random(int(off[-1]))
    flat = np.searchsorted(cdf, u, side="right").astype(np.int32)
    np.minimum(flat, vocab - 1, out=flat)
    return flat, off


def text_corpus_tokens_range(lo: int, hi: int, seed: int = 1235, vocab: int = VOCAB, mean_len: int = 80, min_len: int = 8,
                             chunk: int = 65536):
    """Docs [lo, hi) of a text corpus whose 65536-doc chunks are seeded independently (SeedSequence([seed, chunk_index])),
    so a rank only ever generates its own shard: -> (flat_tokens int32, doc_offsets int64 [hi - lo + 1], shard-local).
    (A different corpus than ``text_corpus_tokens``, same distribution; used for corpora above 2 M docs.)"""
    cdf = _zipf_cdf(vocab)
    flats, lens_all = [], []
    c = lo // chunk
    while c * chunk < hi:
        rng = np.random.default_rng(np.random.SeedSequence([seed, c]))
        lens = np.maximum(min_len, rng.poisson(mean_len, size=chunk)).astype(np.int64)
        off = np.zeros(chunk + 1, dtype=np.int64)
        np.cumsum(lens, out=off[1:])
        flat = np.searchsorted(cdf, rng.random(int(off[-1])), side="right").astype(np.int32)
        np.minimum(flat, vocab - 1, out=flat)
        a, b = max(lo, c * chunk) - c * chunk, min(hi, (c + 1) * chunk) - c * chunk
        flats.append(flat[off[a]:off[b]])
        lens_all.append(lens[a:b])
        c += 1
    lens = np.concatenate(lens_all) if lens_all else np.zeros(0, np.int64)
    out_off = np.zeros(len(lens) + 1, dtype=np.int64)
    np.cumsum(lens, out=out_off[1:])
    return (np.concatenate(flats) if flats else np.zeros(0, np.int32)), out_off


def query_tokens(b: int, seed: int = 4322, vocab: int = VOCAB, length: int = 6) -> np.ndarray:
    rng = np.random.default_rng(seed)
    cdf = _zipf_cdf(vocab)
    flat = np.searchsorted(cdf, rng.random(b * length), side="right")
    return np.minimum(flat, vocab - 1).astype(np.int32).reshape(b, length)


def token_text(tokens) -> str:
    return " ".join(f"w{int(t)}" for t in tokens)


def texts_from_tokens(flat: np.ndarray, off: np.ndarray, lo: int = 0, hi: int | None = None) -> list[str]:
    hi = len(off) - 1 if hi is None else hi
    return [token_text(flat[off[i]:off[i + 1]]) for i in range(lo, hi)]
