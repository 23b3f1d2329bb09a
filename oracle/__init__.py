"""CPU oracle for the retrieve -> fuse -> rerank hot path of chernistry/sentio.

THIS PACKAGE IS TEST INFRASTRUCTURE.  It restates, in NumPy / plain Python, the arithmetic the reference performs on
this path (each function cites the reference file:line it follows).  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import it -- as the checker or the timed CPU baseline, never as
the product.  The product (sentio_b200/) never imports it and fails loudly when the CUDA library is missing.

Pinning (see DESIGN.md "Oracle"): the reference's own tests hold no numeric golden vectors for this path
(SURVEY.md section 4), so the oracle is pinned against outputs of the reference's OWN modules executed in the build
container (tests/golden/make_golden.py imports /root/reference/src/core/retrievers/{hybrid,sparse,scorers}.py and
src/core/rerankers/jina_reranker.py unmodified, with sys.modules stubs for the packages that cannot be installed
offline) and committed as tests/golden/*.json.  Third-party arithmetic that is not under /root/reference:
  * rank-bm25 == 0.2.2 (poetry.lock:4846) -- restated from its published algorithm in oracle/rank_bm25_port.py;
    the wheel is not available offline, so BM25 parity is anchored on the reference call sites
    (src/core/retrievers/sparse.py:88-100,174-198) and on that restatement: "parity pinned to the restatement".
  * Qdrant v1.7.4 cosine search -- restated as exact cosine (oracle/dense.py).
  * Jina hosted reranker -- no local model exists in the reference; the cross-encoder oracle is HF transformers'
    BertForSequenceClassification (library code present in the image) plus our NumPy restatement of it.
"""
