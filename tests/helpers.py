"""Shared helpers for the parity tests."""
from __future__ import annotations

import zlib

import numpy as np


class HashEmbedder:
    """Same deterministic embedder as tests/golden/make_golden.py (must stay in sync with the fixtures)."""

    def __init__(self, dim):
        self.dim = dim

    def embed_sync(self, text):
        rng = np.random.default_rng(zlib.crc32(text.strip().encode("utf-8")))
        v = rng.standard_normal(self.dim)
        return [float(x) for x in (v / np.linalg.norm(v)).astype(np.float32)]

    def embed_many_sync(self, texts):
        return [self.embed_sync(t) for t in texts]


def assert_topk_matches(ids, scores, counts, want_ids, want_scores, rtol=1e-9, tie_eps=1e-12, what=""):
    """ids/ranks identical; where the ORACLE has (near-)exact fp64 ties a permutation inside the tie group is accepted."""
    n = int(counts)
    assert n == len(want_ids), f"{what}: count {n} != {len(want_ids)}"
    got_ids = list(map(int, ids[:n]))
    got_sc = np.asarray(scores[:n], dtype=np.float64)
    want_sc = np.asarray(want_scores, dtype=np.float64)
    assert np.allclose(got_sc, want_sc, rtol=rtol, atol=1e-12), f"{what}: scores differ"
    if got_ids != list(map(int, want_ids)):
        i = 0
        while i < n:
            j = i
            while j + 1 < n and abs(want_sc[j + 1] - want_sc[i]) <= tie_eps * max(1.0, abs(want_sc[i])):
                j += 1
            assert sorted(got_ids[i:j + 1]) == sorted(map(int, want_ids[i:j + 1])), \
                f"{what}: rank {i}..{j}: {got_ids[i:j + 1]} vs {list(want_ids[i:j + 1])}"
            i = j + 1
