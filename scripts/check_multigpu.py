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
    flag = torch.tensor([1 if ok else 0], device=f"cuda:{local}")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
