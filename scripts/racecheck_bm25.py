#!/usr/bin/env python
"""Small multi-range BM25 workload for `compute-sanitizer --tool racecheck`: enough (query, range) items that one CTA walks
several consecutive doc ranges (cursor double buffer, shared-memory accumulators, PLUS bitmap), checked against the oracle."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from oracle.rank_bm25_port import FastBM25
    from sentio_b200 import synth
    from sentio_b200.engine import B200Engine
    from sentio_b200.index import build_bm25_from_token_ids

    eng = B200Engine(0)
    n, V, B, k = 41_000, 2000, 720, 20
    flat, off = synth.text_corpus_tokens(n, vocab=V)
    queries = synth.query_tokens(B, vocab=V)
    for variant in ("okapi", "plus"):
        idx = build_bm25_from_token_ids(flat, off, variant=variant)
        fast = FastBM25(idx.indptr, idx.post_doc, idx.post_tf, idx.doc_len, idx.idf, idx.avgdl, variant)
        eng.load_bm25(idx)
        terms = [idx.term_ids(q) for q in queries]
        ids, sc, cnt = eng.bm25_topk(terms, k)
        for b in range(0, B, 97):
            want = fast.get_scores(list(terms[b]))
            order = np.argsort(-want, kind="stable")[:k]
            order = order[want[order] > 0]
            assert np.array_equal(ids[b, :cnt[b]], order) and np.array_equal(sc[b, :cnt[b]], want[order]), (variant, b)
        print(variant, "ok", flush=True)
    eng.close()


if __name__ == "__main__":
    main()
