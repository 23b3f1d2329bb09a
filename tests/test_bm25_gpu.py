"""K2 parity: BM25 scores bit-identical to the rank_bm25 restatement; top-k identical under the (score desc, index asc)
order; reference-class behaviour (score > 0 filter, unknown tokens, duplicate query tokens, Plus variant)."""
import numpy as np
import pytest

from conftest import load_golden
from oracle.rank_bm25_port import BM25Okapi, BM25Plus, FastBM25
from sentio_b200 import synth
from sentio_b200.document import Document
from sentio_b200.index import build_bm25_from_texts, build_bm25_from_token_ids
from sentio_b200.retrievers.sparse import BM25Retriever

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("variant", ["okapi", "plus"])
def test_scores_bit_exact_vs_faithful_port_on_golden_corpora(engine, variant):
    for c in [x for x in load_golden("bm25") if x["variant"] == variant]:
        idx = build_bm25_from_texts(c["texts"], variant=variant)
        engine.load_bm25(idx)
        for qc in c["queries"]:
            terms = idx.term_ids(qc["query"].lower().split())
            got = engine.bm25_scores(terms)
            assert np.array_equal(got, np.asarray(qc["scores"])), (variant, qc["query"])
            ids, sc, cnt = engine.bm25_topk([terms], 10)
            want = qc["top10"]
            assert int(cnt[0]) == len(want)
            assert [f"d{i}" for i in ids[0, :cnt[0]]] == [w[0] for w in want]
            assert list(sc[0, :cnt[0]]) == [w[1] for w in want]


@pytest.mark.parametrize("variant", ["okapi", "plus"])
def test_retriever_class_matches_golden(variant, monkeypatch):
    monkeypatch.delenv("BM25_VARIANT", raising=False)
    for c in [x for x in load_golden("bm25") if x["variant"] == variant]:
        docs = [Document(id=f"d{i}", text=t) for i, t in enumerate(c["texts"])]
        r = BM25Retriever(documents=docs, variant=variant)
        for qc in c["queries"]:
            out = r.retrieve(qc["query"], top_k=10)
            assert [[d.id, d.metadata["bm25_score"]] for d in out] == qc["top10"]
            assert all(d is r.doc_map[d.id] for d in out)  # shared corpus objects, mutated in place (sparse.py:189-197)


@pytest.mark.parametrize("variant,n", [("okapi", 20000), ("plus", 6000)])
def test_medium_synthetic_corpus_bit_exact_and_topk(engine, variant, n):
    flat, off = synth.text_corpus_tokens(n, vocab=5000)
    idx = build_bm25_from_token_ids(flat, off, variant=variant)
    fast = FastBM25(idx.indptr, idx.post_doc, idx.post_tf, idx.doc_len, idx.idf, idx.avgdl, variant)
    engine.load_bm25(idx, id_base=77)
    queries = synth.query_tokens(24, vocab=5000)
    term_lists = [idx.term_ids(q) for q in queries] + [np.array([-1, -1], np.int32), np.zeros(0, np.int32)]
    k = 100
    ids, sc, cnt = engine.bm25_topk(term_lists, k)
    for b, terms in enumerate(term_lists):
        want = fast.get_scores(list(terms))
        if b < 4:
            assert np.array_equal(engine.bm25_scores(terms), want)
        order = np.argsort(-want, kind="stable")[:k]
        order = order[want[order] > 0]
        assert int(cnt[b]) == len(order), b
        assert np.array_equal(ids[b, :cnt[b]] - 77, order), b
        assert np.array_equal(sc[b, :cnt[b]], want[order]), b


@pytest.mark.parametrize("variant", ["okapi", "plus"])
def test_many_queries_multi_range_ctas(engine, variant):
    """Enough (query, range) work items that one CTA walks SEVERAL consecutive doc ranges and carries the posting
    cursors from one range to the next (ranges_per_cta > 1 in bm25.cu)."""
    n, V, B, k = 100_000, 5000, 320, 50
    flat, off = synth.text_corpus_tokens(n, vocab=V)
    idx = build_bm25_from_token_ids(flat, off, variant=variant)
    fast = FastBM25(idx.indptr, idx.post_doc, idx.post_tf, idx.doc_len, idx.idf, idx.avgdl, variant)
    engine.load_bm25(idx)
    queries = synth.query_tokens(B, vocab=V)
    queries[:, -1] = V - 1 - np.arange(B) % 50  # rare tail terms: posting lists that end long before the last range
    term_lists = [idx.term_ids(q) for q in queries]
    ids, sc, cnt = engine.bm25_topk(term_lists, k)
    for b in range(0, B, 9):
        want = fast.get_scores(list(term_lists[b]))
        order = np.argsort(-want, kind="stable")[:k]
        order = order[want[order] > 0]
        assert int(cnt[b]) == len(order), b
        assert np.array_equal(ids[b, :cnt[b]], order), b
        assert np.array_equal(sc[b, :cnt[b]], want[order]), b


def test_faithful_port_small_random_corpus_okapi_negative_idf(engine):
    rng = np.random.default_rng(4)
    texts = ["common " + " ".join(f"t{rng.integers(0, 12)}" for _ in range(rng.integers(1, 9))) for _ in range(40)]
    ref = BM25Okapi([t.lower().split() for t in texts])
    idx = build_bm25_from_texts(texts)
    assert ref.idf["common"] == idx.idf[idx.vocab["common"]]  # epsilon-floored (negative raw idf)
    engine.load_bm25(idx)
    for q in ["common", "common t1 t1 t2", "t3 zz t4"]:
        toks = q.split()
        assert np.array_equal(engine.bm25_scores(idx.term_ids(toks)), ref.get_scores(toks))


def test_device_entry_point_equals_host_entry_point(engine):
    import torch

    flat, off = synth.text_corpus_tokens(30000, vocab=3000)
    idx = build_bm25_from_token_ids(flat, off)
    engine.load_bm25(idx)
    term_lists = [idx.term_ids(q) for q in synth.query_tokens(70, vocab=3000)]
    h = engine.bm25_topk(term_lists, 50)
    f, o = engine.pack_queries(term_lists)
    d = engine.bm25_topk_dev(torch.from_numpy(f).cuda(), torch.from_numpy(o).cuda(), len(term_lists), int(o[-1]), 6, 50)
    torch.cuda.synchronize()
    for a, b in zip(h, d):
        assert np.array_equal(a, b.cpu().numpy())


def test_full_size_1m_docs_properties(engine):
    """BASELINE config 3 shape (1 M docs, vocab 50 k, Zipf 1.07): exact score spot checks + top-k order properties."""
    n = 1_000_000
    flat, off = synth.text_corpus_tokens(n)
    idx = build_bm25_from_token_ids(flat, off)
    engine.load_bm25(idx)
    fast = FastBM25(idx.indptr, idx.post_doc, idx.post_tf, idx.doc_len, idx.idf, idx.avgdl)
    queries = synth.query_tokens(16)
    term_lists = [idx.term_ids(q) for q in queries]
    ids, sc, cnt = engine.bm25_topk(term_lists, 100)
    for b in range(3):
        want = fast.get_scores(list(term_lists[b]))
        order = np.argsort(-want, kind="stable")[:100]
        order = order[want[order] > 0]
        assert np.array_equal(ids[b, :cnt[b]], order) and np.array_equal(sc[b, :cnt[b]], want[order])
    for b in range(16):
        c = int(cnt[b])
        s = sc[b, :c]
        assert np.all(s > 0) and np.all(np.diff(s) <= 0)
        tie = np.diff(s) == 0
        assert np.all(np.diff(ids[b, :c])[tie] > 0)  # ties ordered by ascending doc index


def test_pyserini_parameters_k1_09_b_04_bit_exact(engine, monkeypatch):
    """SURVEY 8(f)4 / reference sparse.py:219, factory.py:150-157: Pyserini's defaults (k1 = 0.9, b = 0.4) on the GPU
    Okapi kernel, bit-exact against BOTH the faithful rank_bm25 port and the CSR port with the same parameters."""
    monkeypatch.delenv("BM25_VARIANT", raising=False)
    from sentio_b200.retrievers.sparse import PyseriniBM25Retriever

    flat, off = synth.text_corpus_tokens(20000, vocab=3000)
    idx = build_bm25_from_token_ids(flat, off, k1=0.9, b=0.4)
    fast = FastBM25(idx.indptr, idx.post_doc, idx.post_tf, idx.doc_len, idx.idf, idx.avgdl, "okapi", k1=0.9, b=0.4)
    engine.load_bm25(idx)
    term_lists = [idx.term_ids(q) for q in synth.query_tokens(20, vocab=3000)]
    ids, sc, cnt = engine.bm25_topk(term_lists, 30)
    for b, terms in enumerate(term_lists):
        want = fast.get_scores(list(terms))
        assert np.array_equal(engine.bm25_scores(terms), want)
        order = np.argsort(-want, kind="stable")[:30]
        order = order[want[order] > 0]
        assert np.array_equal(ids[b, :cnt[b]], order) and np.array_equal(sc[b, :cnt[b]], want[order])
    # the retriever class with the corpus given explicitly, against the faithful (dict-based) port
    rng = np.random.default_rng(8)
    texts = [" ".join(f"t{rng.integers(0, 60)}" for _ in range(rng.integers(3, 30))) for _ in range(300)]
    ref = BM25Okapi([t.lower().split() for t in texts], k1=0.9, b=0.4)
    r = PyseriniBM25Retriever(documents=[Document(id=f"d{i}", text=t) for i, t in enumerate(texts)])
    for q in ["t1 t2 t3", "t59 t59 t0", "t7"]:
        want = ref.get_scores(q.split())
        order = np.argsort(-want, kind="stable")[:12]
        order = order[want[order] > 0]
        got = r.retrieve(q, top_k=12)
        assert [d.id for d in got] == [f"d{i}" for i in order]
        assert [d.metadata["bm25_score"] for d in got] == [float(want[i]) for i in order]
    with pytest.raises(RuntimeError):   # constructor contract of the reference class: no index directory -> RuntimeError
        PyseriniBM25Retriever(index_dir="/nonexistent/lucene-index")


def test_save_load_round_trip_into_a_fresh_retriever(tmp_path, monkeypatch):
    """Reference sparse.py:102-157: save() -> a FRESH object .load() -> identical retrieve output; a foreign pickle or a
    missing file returns False and leaves the retriever untouched."""
    import pickle

    monkeypatch.delenv("BM25_VARIANT", raising=False)
    rng = np.random.default_rng(3)
    texts = [" ".join(f"w{rng.integers(0, 80)}" for _ in range(rng.integers(4, 40))) for _ in range(500)]
    docs = [Document(id=f"doc-{i}", text=t, metadata={"source": f"s{i % 3}"}) for i, t in enumerate(texts)]
    queries = ["w1 w2 w3", "w79 w0 w0", "unknown-token w5", "w10"]
    for variant in ("okapi", "plus"):
        a = BM25Retriever(documents=docs, variant=variant, cache_dir=str(tmp_path / variant))
        want = [[(d.id, d.metadata["bm25_score"], d.text) for d in a.retrieve(q, top_k=25)] for q in queries]
        a.save()                                                    # default path: <cache_dir>/bm25_index.pkl
        b = BM25Retriever(cache_dir=str(tmp_path / variant))
        assert b.retrieve("w1", top_k=5) == []                      # nothing indexed yet
        assert b.load() is True and b.variant == variant
        got = [[(d.id, d.metadata["bm25_score"], d.text) for d in b.retrieve(q, top_k=25)] for q in queries]
        assert got == want
        assert all(d.metadata["source"] in ("s0", "s1", "s2") for d in b.retrieve(queries[0], top_k=25))  # doc_map restored
        # explicit file path + failure modes
        path = str(tmp_path / f"{variant}.pkl")
        a.save(path)
        c = BM25Retriever()
        assert c.load(str(tmp_path / "missing.pkl")) is False
        foreign = str(tmp_path / "foreign.pkl")
        with open(foreign, "wb") as f:
            pickle.dump({"bm25": object(), "doc_ids": ["x"], "doc_map": {}, "variant": "okapi"}, f)  # reference-style cache
        assert c.load(path) is True
        before = [(d.id, d.metadata["bm25_score"]) for d in c.retrieve(queries[1], top_k=9)]
        assert c.load(foreign) is False
        assert [(d.id, d.metadata["bm25_score"]) for d in c.retrieve(queries[1], top_k=9)] == before == \
               [(i, s) for i, s, _ in want[1][:9]]


def test_top_k_beyond_one_kernel_call(engine, monkeypatch):
    """the reference's argsort[:top_k] supports any top_k; beyond the kernel's 1024 the device still scores
    (sb_bm25_scores) and the retriever applies the reference's cut -- never an empty list."""
    monkeypatch.delenv("BM25_VARIANT", raising=False)
    rng = np.random.default_rng(5)
    texts = [" ".join(f"w{rng.integers(0, 30)}" for _ in range(rng.integers(5, 25))) for _ in range(3000)]
    r = BM25Retriever(documents=[Document(id=f"d{i}", text=t) for i, t in enumerate(texts)])
    ref = BM25Okapi([t.split() for t in texts])
    want = ref.get_scores(["w3", "w4"])
    order = np.argsort(-want, kind="stable")[:2000]
    order = order[want[order] > 0]
    got = r.retrieve("w3 w4", top_k=2000)
    assert len(got) == len(order) > 1024
    assert [d.id for d in got] == [f"d{i}" for i in order]
    assert [d.metadata["bm25_score"] for d in got] == [float(want[i]) for i in order]
    assert len(r.retrieve("w3 w4", top_k=10**6)) == int((want > 0).sum())
