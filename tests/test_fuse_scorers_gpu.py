"""K3 / K4 parity: fusion bit-exact vs the reference's own HybridRetriever outputs (golden) and vs the oracle on random
batches; semantic / MMR scorer signals vs the reference's scorer outputs."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import fusion as fusion_oracle
from oracle import scorers as scorers_oracle

pytestmark = pytest.mark.gpu


def _codes(case):
    code = {}
    for lst in (case["dense"], case["sparse"], case["plugin"]):
        for i, _ in lst:
            code.setdefault(i, len(code))
    return code


def _arr(lst, code):
    if not lst:
        return None
    return (np.asarray([[code[i] for i, _ in lst]], np.int64), np.asarray([[s for _, s in lst]], np.float64),
            np.asarray([len(lst)], np.int32))


def test_fusion_golden_bit_exact(engine):
    for c in load_golden("fusion"):
        code = _codes(c)
        if not code:
            continue
        merged = []
        for i, _ in c["dense"] + c["sparse"]:
            if i not in merged:
                merged.append(i)
        extra = None
        if c["extras"] and merged:
            extra = np.asarray([[[e.get(i, 0.0) for i in merged] for e in c["extras"]]], np.float64)
        ids, sc, src, cnt = engine.fuse(c["method"], c["rrf_k"], c["dense_weight"], c["sparse_weight"], c["top_k"],
                                        dense=_arr(c["dense"], code), sparse=_arr(c["sparse"], code),
                                        plugin=_arr(c["plugin"], code), extra=extra)
        name_of = {v: k for k, v in code.items()}
        got = [[name_of[int(ids[0, j])], float(sc[0, j])] for j in range(int(cnt[0])) if src[0, j] != 0]
        assert got == c["expected"], c["name"]


@pytest.mark.parametrize("method", ["rrf", "weighted_rrf", "comb_sum"])
def test_fusion_random_batches_vs_oracle(engine, method):
    rng = np.random.default_rng(hash(method) % 1000)
    B, k, stride = 37, 100, 100
    d_ids = np.stack([rng.permutation(400)[:stride] for _ in range(B)]).astype(np.int64)
    s_ids = np.stack([rng.permutation(400)[:stride] for _ in range(B)]).astype(np.int64)
    d_sc = -np.sort(-rng.random((B, stride)), axis=1)
    s_sc = -np.sort(-rng.random((B, stride)) * 30, axis=1)
    d_n = rng.integers(0, stride + 1, B).astype(np.int32)
    s_n = rng.integers(0, stride + 1, B).astype(np.int32)
    d_n[0], s_n[0] = stride, stride
    d_n[1], s_n[1] = 0, 5
    ids, sc, src, cnt = engine.fuse(method, 60, 0.7, 0.3, k, dense=(d_ids, d_sc, d_n), sparse=(s_ids, s_sc, s_n))
    for b in range(B):
        d = [(int(d_ids[b, j]), float(d_sc[b, j])) for j in range(d_n[b])]
        s = [(int(s_ids[b, j]), float(s_sc[b, j])) for j in range(s_n[b])]
        want = fusion_oracle.fuse(method, 60, 0.7, 0.3, d, s, [], k)
        assert int(cnt[b]) == len(want)
        assert [int(x) for x in ids[b, :cnt[b]]] == [w[0] for w in want], (method, b)
        assert [float(x) for x in sc[b, :cnt[b]]] == [w[1] for w in want], (method, b)


def test_scorers_golden(engine):
    for c in load_golden("scorers")["semantic_mmr"]:
        q = np.asarray(c["q"], np.float32)
        docs = np.asarray(c["docs"], np.float32)
        sem, mmr = engine.semantic_mmr(q, cand=docs, w_sem=c["w_sem"], lambda_=c["lambda_"], w_mmr=c["w_mmr"])
        # the C ABI takes fp32 embeddings: the survey cases hold fp64 literals (0.9, 0.1, ...) that are not fp32
        # representable, so compare (a) tightly against the oracle on the fp32-rounded inputs and (b) against the
        # reference's own output at the north-star tolerance (1e-3 relative would do; fp32 rounding gives ~1e-7)
        q64, d64 = q.astype(np.float64), [r.astype(np.float64) for r in docs]
        assert np.allclose(sem, scorers_oracle.semantic(q64, d64, c["w_sem"]), rtol=1e-9, atol=1e-12), c["name"]
        assert np.allclose(mmr, scorers_oracle.mmr(q64, d64, c["lambda_"], c["w_mmr"]), rtol=1e-9, atol=1e-12), c["name"]
        assert np.allclose(sem, c["sem"], rtol=1e-6, atol=1e-9), c["name"]
        assert np.allclose(mmr, c["mmr"], rtol=1e-6, atol=1e-9), c["name"]


def test_mmr_300_candidates_1024d_vs_oracle(engine):
    rng = np.random.default_rng(8)
    n, d = 300, 1024
    docs = rng.standard_normal((n, d)).astype(np.float32)
    docs[10:20] = docs[0:10] + 0.05 * rng.standard_normal((10, d)).astype(np.float32)  # near duplicates
    q = rng.standard_normal(d).astype(np.float32)
    sem, mmr = engine.semantic_mmr(q, cand=docs, w_sem=0.8, lambda_=0.5, w_mmr=0.5)
    q64 = q.astype(np.float64)
    d64 = [r.astype(np.float64) for r in docs]
    assert np.allclose(sem, scorers_oracle.semantic(q64, d64, 0.8), rtol=1e-9, atol=1e-12)
    assert np.allclose(mmr, scorers_oracle.mmr(q64, d64, 0.5, 0.5), rtol=1e-9, atol=1e-12)


def test_scorers_from_stored_corpus_vectors(engine):
    rng = np.random.default_rng(12)
    x = rng.standard_normal((500, 128)).astype(np.float32)
    engine.load_dense(x, id_base=10)
    ids = np.asarray([10, 15, 509, 200, 15], np.int64)
    q = rng.standard_normal(128).astype(np.float32)
    sem, mmr = engine.semantic_mmr(q, cand_ids=ids, w_sem=1.0, lambda_=0.7, w_mmr=0.5)
    stored = engine.dense_fetch(ids)
    sem2, mmr2 = engine.semantic_mmr(q, cand=stored, w_sem=1.0, lambda_=0.7, w_mmr=0.5)
    assert np.array_equal(sem, sem2) and np.array_equal(mmr, mmr2)
