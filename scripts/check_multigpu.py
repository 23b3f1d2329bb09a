#!/usr/bin/env python
"""torchrun check of the sharded path on real GPUs: corpus partitioned over WORLD_SIZE ranks, one NCCL all-gather of
per-shard top-k, merge, fusion  ==  the single-GPU result (computed on rank 0 with an unsharded engine).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 scripts/check_multigpu.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist

    from sentio_b200 import synth
    from sentio_b200.index import build_bm25_from_token_ids
    from sentio_b200.pipeline import HybridPipeline

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    n, d, k, B = 300_000, 512, 100, 48
    x = synth.dense_corpus(n, d)
    flat, off = synth.text_corpus_tokens(n, vocab=20000)
    idx = build_bm25_from_token_ids(flat, off)
    q = synth.query_vectors(B, d)
    terms = [idx.term_ids(t) for t in synth.query_tokens(B, vocab=20000)]
    lo, hi = (n * rank) // world, (n * (rank + 1)) // world
    pipe = HybridPipeline(local, rank=rank, world=world)
    pipe.load_dense(x[lo:hi], id_base=lo)
    pipe.load_bm25(idx.shard(lo, hi), id_base=lo)
    d_ids, d_sc, d_cnt = pipe.search_dense(q, k)
    results = {}
    for method in ("rrf", "comb_sum"):
        results[method] = pipe.search_hybrid(q, terms, k, method=method, rrf_k=60, w_dense=0.6, w_sparse=0.4)
    ok = True
    if rank == 0:
        single = HybridPipeline(local)
        single.load_dense(x)
        single.load_bm25(idx)
        rd = single.search_dense(q, k)
        ok &= bool(np.array_equal(rd[0], d_ids) and np.array_equal(rd[1], d_sc) and np.array_equal(rd[2], d_cnt))
        print("dense sharded == single:", ok, flush=True)
        for method in ("rrf", "comb_sum"):
            rs = single.search_hybrid(q, terms, k, method=method, rrf_k=60, w_dense=0.6, w_sparse=0.4)
            same = bool(np.array_equal(rs[0], results[method][0]) and np.array_equal(rs[1], results[method][1]))
            print(f"hybrid {method} sharded == single:", same, flush=True)
            ok &= same
    # the shard built ON THE DEVICE from the local tokens + all-gathered statistics (no global host index) == the shard cut
    # from the global host-built index
    pipe_b = HybridPipeline(local, rank=rank, world=world)
    pipe_b.load_dense(x[lo:hi], id_base=lo)
    sidx = pipe_b.build_bm25_sharded(flat[off[lo]:off[hi]], off[lo:hi + 1] - off[lo], id_base=lo)
    terms_b = [sidx.term_ids(t) for t in synth.query_tokens(B, vocab=20000)]
    rb = pipe_b.search_hybrid(q, terms_b, k, method="rrf", rrf_k=60, w_dense=0.6, w_sparse=0.4)
    same = bool(np.array_equal(rb[0], results["rrf"][0]) and np.array_equal(rb[1], results["rrf"][1]))
    same &= sidx.avgdl == idx.avgdl and sidx.average_idf == idx.average_idf
    print(f"[rank {rank}] device-built shard (all-gathered statistics) == host-built shard:", same, flush=True)
    ok &= same
    pipe_b.engine.close()
    # rerank: every rank holds the merged candidates of all queries and reranks ITS slice of the queries
    from sentio_b200.cross_encoder import CrossEncoderWeights
    from sentio_b200.index import doc_token_matrix, hash_vocab_ids

    cfg = dict(vocab_size=30522, hidden=128, layers=2, heads=4, intermediate=256, max_pos=128, type_vocab=2, ln_eps=1e-12)
    w = CrossEncoderWeights.random(cfg, seed=5, std=0.05)
    vocab_ids = hash_vocab_ids(20000)
    doc_tok, doc_len = doc_token_matrix(flat, off, vocab_ids, ld=120)
    q_raw = synth.query_tokens(B, vocab=20000)
    q_tok = vocab_ids[q_raw].astype(np.int32)
    q_len = np.full(B, q_raw.shape[1], np.int32)
    pipe.load_cross_encoder(w)
    pipe.load_doc_tokens(doc_tok, doc_len)
    r_ids, r_sc, r_cnt = pipe.search_hybrid_rerank(q, terms, q_tok, q_len, 40, 10, seq_len=128)
    a, b = pipe.rerank_slice(B)
    mine = torch.zeros(1, device=f"cuda:{local}")
    if rank == 0:
        single.load_cross_encoder(w)
        single.load_doc_tokens(doc_tok, doc_len)
        s_ids, s_sc, s_cnt = single.search_hybrid_rerank(q, terms, q_tok, q_len, 40, 10, seq_len=128)
        ref = torch.from_numpy(np.concatenate([s_ids.astype(np.float64), s_sc.astype(np.float64)], axis=1)).to(f"cuda:{local}")
    else:
        ref = torch.empty((B, 20), dtype=torch.float64, device=f"cuda:{local}")
    dist.broadcast(ref, 0)
    ref = ref.cpu().numpy()
    same = bool(np.array_equal(ref[a:b, :10], r_ids.astype(np.float64)) and
                np.allclose(ref[a:b, 10:], r_sc.astype(np.float64), rtol=1e-5, atol=1e-6))
    print(f"rank {rank}: rerank slice [{a},{b}) == single-GPU rows: {same}", flush=True)
    mine[0] = 0 if same else 1
    dist.all_reduce(mine)
    if rank == 0:
        ok &= float(mine.item()) == 0
    flag = torch.tensor([1 if ok else 0], device=f"cuda:{local}")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
