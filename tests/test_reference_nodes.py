"""The reference's OWN LangGraph node functions (src/core/graph/nodes.py:37-227), imported unmodified, driving this
repository's retriever / reranker classes -- north_star's acceptance sentence ("so the LangGraph nodes ... call it
unchanged") as a test.

* CPU (`not gpu`): the classes run on the oracle-backed engine double (host logic); skipped when the reference tree is
  not available (neither /root/reference nor the shipped, git-ignored snapshot baseline/_ref).
* GPU: the same nodes on the real engine / C ABI (needs baseline/_ref on the box; `__graft_entry__.build()` takes it).
"""
import threading

import numpy as np
import pytest

from helpers import HashEmbedder
from oracle import refload
from sentio_b200.cross_encoder import CrossEncoderWeights
from sentio_b200.document import Document
from sentio_b200.rerankers.b200_reranker import B200Reranker
from sentio_b200.retrievers.dense import DenseRetriever
from sentio_b200.retrievers.hybrid import HybridRetriever

needs_reference = pytest.mark.skipif(not refload.available(), reason="reference tree not available")

DIM = 48
TEXTS = [f"topic{i % 9} w{i % 13} w{(i * 7) % 31} alpha{i % 5} chunk number {i}" for i in range(240)]
IDS = [f"doc-{i}" for i in range(len(TEXTS))]
QUERIES = ["topic3 w4 alpha2", "w7 chunk", "nothing-in-the-vocabulary", "topic8 topic8 w30"]
CE_CFG = dict(vocab_size=30522, hidden=128, layers=2, heads=4, intermediate=256, max_pos=64, type_vocab=2, ln_eps=1e-12)


def _build(make_store, make_sparse, engine):
    emb = HashEmbedder(DIM)
    vecs = np.asarray(emb.embed_many_sync(TEXTS), dtype=np.float32)
    payloads = [{"content": t, "metadata": {"source": f"s{i % 4}", "page": i}} for i, t in enumerate(TEXTS)]
    store = make_store(vecs, IDS, payloads)
    corpus = [Document(id=i, text=t, metadata={"source": "corpus"}) for i, t in zip(IDS, TEXTS)]
    dense = DenseRetriever(client=store, embedder=emb, collection_name="Sentio_docs")
    hr = HybridRetriever(dense_retriever=dense, sparse_retriever=make_sparse(corpus), rrf_k=60, scorer_plugins=[],
                         fusion_method="rrf", engine=engine)
    rr = B200Reranker(weights=CrossEncoderWeights.random(CE_CFG, seed=3), engine=engine, seq_len=48)
    return hr, rr


def _drive_nodes(hr, rr):
    ref = refload.load()
    assert hasattr(ref, "create_retriever_node"), getattr(ref, "graph_import_error", None)
    retrieve_node = ref.create_retriever_node(hr, top_k=10)
    rerank_node = ref.create_reranker_node(rr, top_k=4)
    for q in QUERIES:
        # ---- retrieve_node == HybridRetriever.retrieve (nodes.py:51-119)
        state = retrieve_node(ref.create_initial_state(q))
        want = hr.retrieve(q, top_k=10)
        got = state["retrieved_documents"]
        assert "retriever_error" not in state["metadata"], state["metadata"]
        assert [d.id for d in got] == [d.id for d in want]
        assert [d.metadata["score"] for d in got] == [d.metadata["score"] for d in want]
        assert [d.metadata["hybrid_score"] for d in got] == [d.metadata["hybrid_score"] for d in want]
        assert all(type(d) is ref.Document for d in got)          # the node re-wraps into the reference's dataclass
        assert [d.text for d in got] == [d.text for d in want] and all(d.text for d in got)
        assert state["metadata"]["retriever_type"] == "HybridRetriever"
        assert state["metadata"]["retrieved_count"] == len(want)
        # ---- metadata.user_top_k overrides the node's top_k (nodes.py:64-69)
        st5 = ref.create_initial_state(q)
        st5["metadata"]["user_top_k"] = 5
        want5 = hr.retrieve(q, top_k=5)   # (a hybrid top-5 is not a prefix of the top-10: the sub-retrievers get top_k too)
        assert [d.id for d in retrieve_node(st5)["retrieved_documents"]] == [d.id for d in want5] and len(want5) <= 5
        # ---- rerank_node == B200Reranker.rerank on the node's prepared copies (nodes.py:138-227)
        direct = rr.rerank(query=q, docs=[Document(id=d.id, text=d.text, metadata=dict(d.metadata)) for d in got], top_k=4)
        state = rerank_node(state)
        rer = state["reranked_documents"]
        if not got:
            assert rer == []
            continue
        assert "reranker_error" not in state["metadata"], state["metadata"]
        assert [d.id for d in rer] == [d.id for d in direct]
        assert [d.metadata["rerank_score"] for d in rer] == [d.metadata["rerank_score"] for d in direct]
        assert all(0.0 <= d.metadata["score"] <= 1.0 and d.metadata["score"] == d.metadata["rerank_score"] for d in rer)
        assert state["metadata"]["reranker_type"] == "B200Reranker" and state["metadata"]["reranked_count"] == len(rer)
        sc = [d.metadata["rerank_score"] for d in rer]
        assert sc == sorted(sc, reverse=True)

    # ---- a raising retriever lands in metadata["retriever_error"], the graph continues without documents
    class Boom:
        def retrieve(self, query, top_k=10):
            raise RuntimeError("index offline")

    st = ref.create_retriever_node(Boom(), top_k=3)(ref.create_initial_state("q"))
    assert st["metadata"]["retriever_error"] == "index offline" and st["retrieved_documents"] == []
    # ---- no documents: rerank_node returns the state untouched
    st = rerank_node(ref.create_initial_state("q"))
    assert st["reranked_documents"] == [] and "reranker_type" not in st["metadata"]


@needs_reference
def test_reference_nodes_drive_the_repo_classes_host_logic(monkeypatch):
    from oracle_engine import OracleEngine
    from sentio_b200.retrievers import sparse as sparse_mod
    from test_hybrid_e2e import _OracleStore

    monkeypatch.delenv("BM25_VARIANT", raising=False)
    monkeypatch.setattr(sparse_mod, "B200Engine", lambda device=0: OracleEngine())
    eng = OracleEngine()
    hr, rr = _build(_OracleStore, lambda corpus: sparse_mod.BM25Retriever(documents=corpus), eng)
    _drive_nodes(hr, rr)


@pytest.mark.gpu
@needs_reference
def test_reference_nodes_drive_the_repo_classes_on_the_gpu(engine, monkeypatch):
    from sentio_b200.retrievers.sparse import BM25Retriever
    from sentio_b200.vector_store import B200VectorStore

    monkeypatch.delenv("BM25_VARIANT", raising=False)

    def make_store(vecs, ids, payloads):
        st = B200VectorStore(0)
        st.create_collection("Sentio_docs", vecs, ids=ids, payloads=payloads)
        return st

    hr, rr = _build(make_store, lambda corpus: BM25Retriever(documents=corpus), engine)
    _drive_nodes(hr, rr)


@pytest.mark.gpu
def test_concurrent_retrieve_async_on_one_context(engine, monkeypatch):
    """retrievers/base.py:37-42 dispatches ``retrieve`` to the default thread pool, so one engine context is entered
    from several Python threads at once (ctypes releases the GIL): every call must return exactly the serial answer."""
    import asyncio

    from sentio_b200.retrievers.sparse import BM25Retriever
    from sentio_b200.vector_store import B200VectorStore

    monkeypatch.delenv("BM25_VARIANT", raising=False)

    def make_store(vecs, ids, payloads):
        st = B200VectorStore(0)
        st.create_collection("Sentio_docs", vecs, ids=ids, payloads=payloads)
        return st

    hr, rr = _build(make_store, lambda corpus: BM25Retriever(documents=corpus), engine)
    queries = [f"topic{i % 9} w{i % 13} alpha{i % 5}" for i in range(48)]
    # ids only: like the reference, BM25Retriever hands out the SHARED corpus Documents and the fusion writes
    # metadata["hybrid_score"] into them in place (sparse.py:189-197, hybrid.py:296-297), so under concurrency a document's
    # score field belongs to whichever query wrote last -- the ranking of each call is what must be stable
    serial = [[d.id for d in hr.retrieve(q, top_k=10)] for q in queries]

    async def one(q):
        docs = await hr.retrieve_async(q, top_k=10)
        return [d.id for d in docs]

    async def hammer():
        return await asyncio.gather(*[one(q) for q in queries])

    assert asyncio.run(hammer()) == serial
    # raw threads on the engine itself: dense / BM25 / rerank entry points interleaved on ONE sb_ctx
    emb = HashEmbedder(DIM)
    qv = np.asarray(emb.embed_many_sync(queries), dtype=np.float32)
    st = hr._dense._client  # the B200VectorStore behind the dense retriever
    want = [st.search("Sentio_docs", list(v), limit=7) for v in qv]
    errors, got = [], [None] * len(queries)

    def worker(lo, hi):
        try:
            for i in range(lo, hi):
                got[i] = st.search("Sentio_docs", list(qv[i]), limit=7)
                rr.score_pairs(queries[i], TEXTS[i:i + 3])
        except Exception as exc:  # pragma: no cover
            errors.append(exc)

    threads = [threading.Thread(target=worker, args=(j * 12, (j + 1) * 12)) for j in range(4)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    assert not errors, errors
    assert [[(p.id, p.score) for p in r] for r in got] == [[(p.id, p.score) for p in r] for r in want]
