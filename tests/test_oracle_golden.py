"""Pin the oracle: every restatement in oracle/ must reproduce the golden vectors that were produced by executing the
reference's own modules (tests/golden/make_golden.py), bit for bit where the arithmetic is the same sequence of fp64 ops."""
import math

import numpy as np
import pytest

from conftest import load_golden
from oracle import fusion as fusion_oracle
from oracle import scorers as scorers_oracle
from oracle.rank_bm25_port import BM25Okapi, BM25Plus, FastBM25
from sentio_b200.index import build_bm25_from_texts, build_bm25_from_token_ids


def test_fusion_oracle_reproduces_reference_bit_exact():
    for c in load_golden("fusion"):
        merged = []
        seen = set()
        for i, _ in c["dense"] + c["sparse"]:
            if i not in seen:
                seen.add(i)
                merged.append(i)
        extras = [[e.get(i, 0.0) for i in merged] for e in c["extras"]]
        out = fusion_oracle.fuse(c["method"], c["rrf_k"], c["dense_weight"], c["sparse_weight"],
                                 [tuple(x) for x in c["dense"]], [tuple(x) for x in c["sparse"]],
                                 [tuple(x) for x in c["plugin"]], c["top_k"], extras)
        got = [[i, s] for i, s, has_doc in out if has_doc]
        assert got == c["expected"], c["name"]


def test_survey_known_answers():
    cases = {c["name"]: c for c in load_golden("fusion")}
    assert cases["survey_rrf"]["expected"] == [["B", 0.03278688524590164], ["A", 0.016666666666666666],
                                                ["D", 0.016666666666666666], ["C", 0.016129032258064516],
                                                ["E", 0.016129032258064516]]
    assert cases["survey_comb"]["expected"][0] == ["A", 0.7]
    sc = {c["name"]: c for c in load_golden("scorers")["semantic_mmr"]}
    assert sc["survey_l05"]["mmr"] == [0.24847093366840473, 0.0, 0.0, 0.0]
    assert sc["survey_l07"]["sem"] == [0.7951069877388952, 0.7704694597489229, 0.0, 0.7938976988447587]
    assert load_golden("scorers")["keyword"]["expected"] == [0.15000000000000002, 0.05]


def test_scorer_oracles_reproduce_reference():
    for c in load_golden("scorers")["semantic_mmr"]:
        q = np.asarray(c["q"])
        docs = [np.asarray(d) for d in c["docs"]]
        assert scorers_oracle.semantic(q, docs, c["w_sem"]) == c["sem"], c["name"]
        assert scorers_oracle.mmr(q, docs, c["lambda_"], c["w_mmr"]) == c["mmr"], c["name"]


@pytest.mark.parametrize("variant", ["okapi", "plus"])
def test_bm25_index_builder_and_fast_oracle_match_rank_bm25(variant):
    cls = BM25Okapi if variant == "okapi" else BM25Plus
    for c in [x for x in load_golden("bm25") if x["variant"] == variant]:
        tokenized = [t.lower().split() for t in c["texts"]]
        ref = cls(tokenized)
        idx = build_bm25_from_texts(c["texts"], variant=variant)
        # vocabulary order = rank_bm25's dict insertion order; idf / avgdl bit-identical
        assert list(idx.vocab.keys()) == list(ref.idf.keys())
        assert idx.avgdl == ref.avgdl == c["avgdl"]
        assert [float(x) for x in idx.idf] == [ref.idf[w] for w in idx.vocab]
        assert {w: float(idx.idf[i]) for w, i in idx.vocab.items()} == c["idf"]
        fast = FastBM25(idx.indptr, idx.post_doc, idx.post_tf, idx.doc_len, idx.idf, idx.avgdl, variant)
        for qc in c["queries"]:
            toks = qc["query"].lower().split()
            want = np.asarray(qc["scores"])
            assert np.array_equal(ref.get_scores(toks), want)
            got = fast.get_scores(list(idx.term_ids(toks)))
            assert np.array_equal(got, want), (variant, qc["query"])


def test_reference_topk_agrees_with_stable_order_up_to_ties():
    """The documented deviation: stable tie order.  The reference's own (unstable argsort) output has the same scores."""
    for c in load_golden("bm25"):
        for qc in c["queries"]:
            assert [s for _, s in qc["top10"]] == [s for _, s in qc["reference_top10"]]
            assert {i for i, _ in qc["top10"]} == {i for i, _ in qc["reference_top10"]} or \
                len(set(s for _, s in qc["top10"])) < len(qc["top10"])


def test_token_id_builder_equals_text_builder():
    rng = np.random.default_rng(0)
    docs = [rng.integers(0, 50, size=int(rng.integers(1, 30))) for _ in range(80)]
    texts = [" ".join(f"w{t}" for t in d) for d in docs]
    a = build_bm25_from_texts(texts)
    flat = np.concatenate(docs)
    off = np.concatenate([[0], np.cumsum([len(d) for d in docs])])
    b = build_bm25_from_token_ids(flat, off)
    assert np.array_equal(a.indptr, b.indptr) and np.array_equal(a.post_doc, b.post_doc)
    assert np.array_equal(a.post_tf, b.post_tf) and np.array_equal(a.idf, b.idf) and a.avgdl == b.avgdl
    q = [3, 3, 7, 49, 1000]
    assert list(b.term_ids(q)) == list(a.term_ids([f"w{t}" for t in q]))


def test_shard_preserves_scores():
    rng = np.random.default_rng(1)
    docs = [rng.integers(0, 30, size=int(rng.integers(1, 20))) for _ in range(101)]
    flat = np.concatenate(docs)
    off = np.concatenate([[0], np.cumsum([len(d) for d in docs])])
    full = build_bm25_from_token_ids(flat, off)
    q = list(full.term_ids([1, 2, 2, 5]))
    want = FastBM25(full.indptr, full.post_doc, full.post_tf, full.doc_len, full.idf, full.avgdl).get_scores(q)
    parts = []
    for lo, hi in [(0, 40), (40, 41), (41, 101)]:
        sh = full.shard(lo, hi)
        parts.append(FastBM25(sh.indptr, sh.post_doc, sh.post_tf, sh.doc_len, sh.idf, sh.avgdl).get_scores(q))
    assert np.array_equal(np.concatenate(parts), want)


def test_cross_encoder_numpy_oracle_matches_hf():
    from oracle import cross_encoder as ce
    from sentio_b200.cross_encoder import CrossEncoderWeights
    from sentio_b200.index import hash_tokenize_pairs

    cfg = dict(vocab_size=30522, hidden=64, layers=2, heads=4, intermediate=128, max_pos=64, type_vocab=2, ln_eps=1e-12)
    model = ce.hf_model(cfg, seed=0)
    w = CrossEncoderWeights.from_hf_state_dict(model.state_dict(), cfg)
    ids, tt, lens = hash_tokenize_pairs("what is retrieval", ["retrieval is search", "", "a b c d e f g h i j k"], 32)
    l_hf, s_hf = ce.hf_scores(model, ids, tt, lens)
    l_np, s_np = ce.numpy_forward(w, ids, tt, lens)
    assert np.allclose(l_np, l_hf, rtol=1e-4, atol=1e-5)
    assert np.allclose(s_np, s_hf, rtol=1e-5)
    assert w.blob().size == sum(int(np.prod(s)) for _, s in CrossEncoderWeights.tensor_order(cfg))


def test_selector_node_matches_reference_golden():
    """§8f row 3: sentio_b200.selector.create_document_selector_node == the reference node (nodes.py:231-372) on the
    committed fixture (generated by tests/golden/make_golden.py from the reference's own code)."""
    from sentio_b200.document import Document
    from sentio_b200.selector import create_document_selector_node

    for c in load_golden("selector"):
        mk = lambda d: Document(id=d["id"], text=d["text"], metadata=dict(d["metadata"]))
        state = dict(query="q", retrieved_documents=[mk(d) for d in c["docs"]], reranked_documents=[],
                     selected_documents=[], response="", metadata={}, evaluation={})
        if c["use_reranked"]:
            state["reranked_documents"] = [mk(d) for d in reversed(c["docs"])]
        if c["user_top_k"] is not None:
            state["metadata"]["user_top_k"] = c["user_top_k"]
        out = create_document_selector_node(top_k=c["top_k"], max_tokens=c["max_tokens"])(state)
        assert [[d.id, d.text, d.metadata] for d in out["selected_documents"]] == c["selected"]
        assert {k: v for k, v in out["metadata"].items() if k != "user_top_k"} == c["meta"]
