"""Whole-stack parity against the reference's own classes (golden hybrid_e2e.json = reference DenseRetriever +
BM25Retriever + HybridRetriever + scorer plugins executed in the build container).

* CPU (`not gpu`): our retriever classes driven by the oracle-backed engine double -> pins the HOST logic.
* GPU: the same classes on the real engine / C ABI -> the parity test proper.
"""
import numpy as np
import pytest

from conftest import load_golden
from helpers import HashEmbedder
from sentio_b200.document import Document
from sentio_b200.retrievers.dense import DenseRetriever
from sentio_b200.retrievers.hybrid import HybridRetriever
from sentio_b200.retrievers.scorers import KeywordMatchScorer, MMRScorer, SemanticSimilarityScorer


def _run(gold, make_store, make_sparse, fusion_engine, scorer_engine):
    emb = HashEmbedder(gold["dim"])
    vecs = np.asarray([emb.embed_sync(t) for t in gold["texts"]], dtype=np.float32)
    payloads = [{"content": t, "metadata": {"source": f"s{i % 7}"}} for i, t in enumerate(gold["texts"])]
    store = make_store(vecs, gold["ids"], payloads)
    for run in gold["runs"]:
        corpus = [Document(id=i, text=t, metadata={"source": "corpus"}) for i, t in zip(gold["ids"], gold["texts"])]
        dense = DenseRetriever(client=store, embedder=emb, collection_name="Sentio_docs")
        sparse = make_sparse(corpus)
        plugins = None
        if run["plugins"]:
            plugins = [SemanticSimilarityScorer(embedder=emb, weight=0.8, engine=scorer_engine),
                       KeywordMatchScorer(weight=0.2),
                       MMRScorer(embedder=emb, lambda_=0.5, weight=0.5, engine=scorer_engine)]
        hr = HybridRetriever(dense_retriever=dense, sparse_retriever=sparse, rrf_k=60, scorer_plugins=plugins,
                             fusion_method=run["method"], dense_weight=0.6, sparse_weight=0.4, engine=fusion_engine)
        for q, want in zip(gold["queries"], run["results"]):
            got = hr.retrieve(q, top_k=15)
            assert [d.id for d in got] == [w[0] for w in want], (run["method"], run["plugins"], q)
            gs = np.asarray([d.metadata["score"] for d in got])
            ws = np.asarray([w[1] for w in want])
            if run["plugins"] or run["method"] == "comb_sum":
                # comb_sum consumes the raw dense cosines, which agree with NumPy's to ~1e-16 (fp64 summation order)
                assert np.allclose(gs, ws, rtol=1e-9, atol=1e-12)
            else:
                assert np.array_equal(gs, ws), (run["method"], q)  # rank fusion: bit-exact
            assert all(d.metadata["hybrid_score"] == d.metadata["score"] for d in got)


class _OracleStore:
    """QdrantClient-shaped store on the oracle engine double (CPU host-logic test only)."""

    def __init__(self, vecs=None, ids=None, payloads=None):
        from sentio_b200.vector_store import ScoredPoint

        self.cols, self.SP = {}, ScoredPoint
        if vecs is not None:
            self.create_collection("Sentio_docs", vecs, ids=ids, payloads=payloads)

    def create_collection(self, name, vecs, ids=None, payloads=None):
        from oracle_engine import OracleEngine

        eng = OracleEngine()
        eng.load_dense(vecs)
        self.cols[name] = (eng, ids, payloads)

    def collection_exists(self, collection_name):
        return collection_name in self.cols

    def search(self, collection_name, query_vector, limit=10, with_payload=True, with_vectors=False):
        eng, ids, payloads = self.cols[collection_name]
        i, s, c = eng.dense_topk(np.asarray(query_vector, np.float32)[None], limit)
        return [self.SP(id=ids[int(i[0, j])], score=float(s[0, j]), payload=payloads[int(i[0, j])])
                for j in range(int(c[0]))]


def _run_cache(gold, make_store, make_sparse, fusion_engine, scorer_engine):
    """golden hybrid_cache.json: the reference stack with a populated ``web_cache`` collection (hybrid.py:146-182,208)."""
    emb = HashEmbedder(gold["dim"])
    store = make_store()
    store.create_collection("Sentio_docs", np.asarray(emb.embed_many_sync(gold["texts"]), np.float32), ids=gold["ids"],
                            payloads=[{"content": t, "metadata": {"source": f"s{i % 5}"}} for i, t in enumerate(gold["texts"])])
    store.create_collection("web_cache", np.asarray(emb.embed_many_sync(gold["cache_texts"]), np.float32),
                            ids=gold["cache_ids"],
                            payloads=[{"content": t, "metadata": {"source": "web"}} for t in gold["cache_texts"]])
    seen_web = seen_both = 0
    for run in gold["runs"]:
        corpus = [Document(id=i, text=t, metadata={"source": "corpus"}) for i, t in zip(gold["ids"], gold["texts"])]
        dense = DenseRetriever(client=store, embedder=emb, collection_name="Sentio_docs")
        plugins = None
        if run["plugins"]:
            plugins = [SemanticSimilarityScorer(embedder=emb, weight=0.8, engine=scorer_engine),
                       KeywordMatchScorer(weight=0.2),
                       MMRScorer(embedder=emb, lambda_=0.5, weight=0.5, engine=scorer_engine)]
        hr = HybridRetriever(dense_retriever=dense, sparse_retriever=make_sparse(corpus), rrf_k=60, scorer_plugins=plugins,
                             fusion_method=run["method"], dense_weight=0.6, sparse_weight=0.4, engine=fusion_engine)
        assert hr._has_cache_collection
        for q, want in zip(gold["queries"], run["results"]):
            got = hr.retrieve(q, top_k=12)
            assert [d.id for d in got] == [w[0] for w in want], (run["method"], run["plugins"], q)
            assert [d.text for d in got] == [w[2] for w in want]       # the cached (edited) text wins, like the reference
            gs, ws = np.asarray([d.metadata["score"] for d in got]), np.asarray([w[1] for w in want])
            if run["plugins"] or run["method"] == "comb_sum":
                assert np.allclose(gs, ws, rtol=1e-9, atol=1e-12)
            else:
                assert np.array_equal(gs, ws), (run["method"], q)
            seen_web += sum(d.id.startswith("web-") for d in got)
            seen_both += sum(d.id in set(gold["cache_ids"]) and d.id.startswith("doc-") for d in got)
    assert seen_web > 0 and seen_both > 0   # the fixture really exercises cache-only and doubly-listed ids


def test_host_logic_with_oracle_engine(monkeypatch):
    from oracle_engine import OracleEngine
    from sentio_b200.retrievers import sparse as sparse_mod

    monkeypatch.delenv("BM25_VARIANT", raising=False)
    monkeypatch.setenv("CACHE_COLLECTION_NAME", "web_cache")
    monkeypatch.setattr(sparse_mod, "B200Engine", lambda device=0: OracleEngine())
    eng = OracleEngine()
    _run(load_golden("hybrid_e2e"), _OracleStore, lambda corpus: sparse_mod.BM25Retriever(documents=corpus), eng, eng)


def test_web_cache_collection_host_logic_with_oracle_engine(monkeypatch):
    from oracle_engine import OracleEngine
    from sentio_b200.retrievers import sparse as sparse_mod

    monkeypatch.delenv("BM25_VARIANT", raising=False)
    monkeypatch.setenv("CACHE_COLLECTION_NAME", "web_cache")
    monkeypatch.setattr(sparse_mod, "B200Engine", lambda device=0: OracleEngine())
    eng = OracleEngine()
    _run_cache(load_golden("hybrid_cache"), _OracleStore, lambda corpus: sparse_mod.BM25Retriever(documents=corpus), eng, eng)


@pytest.mark.gpu
def test_gpu_stack_with_web_cache_collection_matches_reference(engine, monkeypatch):
    from sentio_b200.retrievers.sparse import BM25Retriever
    from sentio_b200.vector_store import B200VectorStore

    monkeypatch.delenv("BM25_VARIANT", raising=False)
    monkeypatch.setenv("CACHE_COLLECTION_NAME", "web_cache")
    _run_cache(load_golden("hybrid_cache"), lambda: B200VectorStore(0), lambda corpus: BM25Retriever(documents=corpus),
               engine, engine)


@pytest.mark.gpu
def test_gpu_stack_matches_reference(engine, monkeypatch):
    from sentio_b200.retrievers.sparse import BM25Retriever
    from sentio_b200.vector_store import B200VectorStore

    monkeypatch.delenv("BM25_VARIANT", raising=False)

    def make_store(vecs, ids, payloads):
        st = B200VectorStore(0)
        st.create_collection("Sentio_docs", vecs, ids=ids, payloads=payloads)
        return st

    _run(load_golden("hybrid_e2e"), make_store, lambda corpus: BM25Retriever(documents=corpus), engine, engine)


@pytest.mark.gpu
def test_pipeline_batch_equals_per_query_classes(engine):
    """HybridPipeline -- the host entry point (sb_hybrid_topk) AND the device-resident batch path -- == dense_topk +
    bm25_topk + fuse composed per stage."""
    from sentio_b200 import synth
    from sentio_b200.index import build_bm25_from_token_ids
    from sentio_b200.pipeline import HybridPipeline

    import torch

    n, d, k, B = 30000, 256, 100, 70
    x = synth.dense_corpus(n, d)
    flat, off = synth.text_corpus_tokens(n, vocab=4000)
    idx = build_bm25_from_token_ids(flat, off)
    pipe = HybridPipeline(0)
    pipe.load_dense(x)
    pipe.load_bm25(idx)
    q = synth.query_vectors(B, d)
    terms = [idx.term_ids(t) for t in synth.query_tokens(B, vocab=4000)]
    terms[5] = np.zeros(0, np.int32)          # a query without text
    terms[6] = np.array([-1, -1], np.int32)   # only unknown tokens
    for method in ("rrf", "comb_sum"):
        ids, sc, src, cnt = pipe.search_hybrid(q, terms, k, method=method, rrf_k=60, w_dense=0.6, w_sparse=0.4)
        fl, of = pipe.engine.pack_queries(terms)
        dev = pipe.hybrid_dev(torch.from_numpy(q).cuda(), torch.from_numpy(fl).cuda(), torch.from_numpy(of).cuda(),
                              int(of[-1]), int(np.diff(of).max()), k, method, 60, 0.6, 0.4)
        torch.cuda.synchronize()
        assert np.array_equal(ids, dev[0].cpu().numpy()) and np.array_equal(sc, dev[1].cpu().numpy())
        assert np.array_equal(src, dev[2].cpu().numpy()) and np.array_equal(cnt, dev[3].cpu().numpy())
        engine.load_dense(x)
        engine.load_bm25(idx)
        dl = engine.dense_topk(q, k)
        sl = engine.bm25_topk(terms, k)
        f = engine.fuse(method, 60, 0.6, 0.4, k, dense=dl, sparse=sl)
        assert np.array_equal(ids, f[0]) and np.array_equal(sc, f[1]) and np.array_equal(cnt, f[3])


def test_batch_retrieval_equals_per_query_with_oracle_engine(monkeypatch):
    """retrieve_batch of the dense / BM25 retrievers == retrieve per query (host logic; oracle-backed engine double)."""
    from helpers import HashEmbedder
    from oracle_engine import OracleEngine
    from sentio_b200 import vector_store as vs_mod
    from sentio_b200.retrievers import sparse as sparse_mod
    from sentio_b200.retrievers.dense import DenseRetriever

    monkeypatch.delenv("BM25_VARIANT", raising=False)
    monkeypatch.setattr(sparse_mod, "B200Engine", lambda device=0: OracleEngine())
    monkeypatch.setattr(vs_mod, "B200Engine", lambda device=0: OracleEngine())
    texts = [f"w{i % 7} w{i % 11} w{i % 13} topic{i % 5}" for i in range(60)]
    docs = [Document(id=f"d{i}", text=t) for i, t in enumerate(texts)]
    emb = HashEmbedder(32)
    store = vs_mod.B200VectorStore(device=0)
    store.create_collection("c", np.asarray(emb.embed_many_sync(texts), np.float32).astype(np.float16),
                            ids=[d.id for d in docs], payloads=[{"content": t} for t in texts])
    dense = DenseRetriever(client=store, embedder=emb, collection_name="c")
    bm25 = sparse_mod.BM25Retriever(documents=docs)
    queries = ["w1 w2 topic3", "w5", "nothing-known", "w6 w6 w10"]
    for r in (dense, bm25):
        key = "score" if r is dense else "bm25_score"
        batch = [[(d.id, d.metadata[key], d.text) for d in x] for x in r.retrieve_batch(queries, top_k=7)]
        for q, got in zip(queries, batch):  # compare query by query: ``retrieve`` mutates the shared corpus objects
            assert got == [(d.id, d.metadata[key], d.text) for d in r.retrieve(q, top_k=7)]
    # batch hits are copies: a document that is a hit of two queries keeps both scores
    hits = bm25.retrieve_batch(["w1", "w1 w2"], top_k=60)
    shared = {d.id for d in hits[0]} & {d.id for d in hits[1]}
    assert shared and all(a is not b for a in hits[0] for b in hits[1] if a.id == b.id)
    assert dense.retrieve_batch([], top_k=3) == [] and bm25.retrieve_batch([], top_k=3) == []
    # HybridRetriever: batched retrieval stages + the same fusion code as the single-query path
    for method in ("rrf", "weighted_rrf", "comb_sum"):
        hr = HybridRetriever(dense_retriever=dense, sparse_retriever=bm25, rrf_k=60, fusion_method=method,
                             dense_weight=0.6, sparse_weight=0.4, scorer_plugins=[], engine=OracleEngine())
        got = [[(d.id, d.metadata["hybrid_score"]) for d in x] for x in hr.retrieve_batch(queries, top_k=9)]
        for q, g in zip(queries, got):
            assert g == [(d.id, d.metadata["hybrid_score"]) for d in hr.retrieve(q, top_k=9)], (method, q)
        assert hr.retrieve_batch([], top_k=9) == []
