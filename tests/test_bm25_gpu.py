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
