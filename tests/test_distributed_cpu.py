"""world_size-2 gloo test of the sharded path (HybridPipeline): corpus partitioned by contiguous doc range, per-shard
top-k, ONE all-gather of the packed records, shard merge, fusion on GLOBAL ranks == the unsharded result.
Runs on CPU with the oracle-backed engine double (the GPU arithmetic has its own parity tests)."""
import os
import socket
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, tmpdir):
    sys.path.insert(0, os.path.dirname(HERE))
    sys.path.insert(0, HERE)
    import torch.distributed as dist

    from oracle_engine import OracleEngineTorch
    from sentio_b200 import synth
    from sentio_b200.index import build_bm25_from_token_ids
    from sentio_b200.pipeline import HybridPipeline

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n, d, k, B = 3001, 32, 20, 7
    x = synth.dense_corpus(n, d)
    flat, off = synth.text_corpus_tokens(n, vocab=300)
    idx = build_bm25_from_token_ids(flat, off)
    q = synth.query_vectors(B, d)
    terms = [idx.term_ids(t) for t in synth.query_tokens(B, vocab=300)]
    lo, hi = (n * rank) // world, (n * (rank + 1)) // world
    pipe = HybridPipeline(device=None, rank=rank, world=world, engine=OracleEngineTorch())
    pipe.load_dense(x[lo:hi], id_base=lo)
    pipe.load_bm25(idx.shard(lo, hi), id_base=lo)
    d_ids, d_sc, d_cnt = pipe.search_dense(q, k)
    f_ids, f_sc, f_src, f_cnt = pipe.search_hybrid(q, terms, k, method="rrf", rrf_k=60)
    c_ids, c_sc, _, c_cnt = pipe.search_hybrid(q, terms, k, method="comb_sum", rrf_k=60, w_dense=0.7, w_sparse=0.3)
    # the same shard built WITHOUT a global host index: local postings + all-gathered statistics (build_bm25_sharded)
    pipe2 = HybridPipeline(device=None, rank=rank, world=world, engine=OracleEngineTorch())
    pipe2.load_dense(x[lo:hi], id_base=lo)
    sidx = pipe2.build_bm25_sharded(flat[off[lo]:off[hi]], off[lo:hi + 1] - off[lo], id_base=lo)
    assert sidx.avgdl == idx.avgdl and sidx.average_idf == idx.average_idf
    raw = np.nonzero(sidx.token_id_map >= 0)[0]
    assert np.array_equal(sidx.idf[sidx.token_id_map[raw]], idx.idf[idx.token_id_map[raw]])   # bit-identical global idf
    terms2 = [sidx.term_ids(t) for t in synth.query_tokens(B, vocab=300)]
    g_ids, g_sc, _, g_cnt = pipe2.search_hybrid(q, terms2, k, method="rrf", rrf_k=60)
    np.savez(os.path.join(tmpdir, f"rank{rank}.npz"), d_ids=d_ids, d_sc=d_sc, d_cnt=d_cnt, f_ids=f_ids, f_sc=f_sc,
             f_cnt=f_cnt, c_ids=c_ids, c_sc=c_sc, c_cnt=c_cnt, g_ids=g_ids, g_sc=g_sc, g_cnt=g_cnt)
    dist.barrier()
    dist.destroy_process_group()


def test_two_shards_equal_unsharded(tmp_path):
    import torch.multiprocessing as mp

    sys.path.insert(0, HERE)
    from oracle_engine import OracleEngineTorch
    from sentio_b200 import synth
    from sentio_b200.index import build_bm25_from_token_ids
    from sentio_b200.pipeline import HybridPipeline

    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    n, d, k, B = 3001, 32, 20, 7
    x = synth.dense_corpus(n, d)
    flat, off = synth.text_corpus_tokens(n, vocab=300)
    idx = build_bm25_from_token_ids(flat, off)
    q = synth.query_vectors(B, d)
    terms = [idx.term_ids(t) for t in synth.query_tokens(B, vocab=300)]
    single = HybridPipeline(device=None, engine=OracleEngineTorch())
    single.load_dense(x)
    single.load_bm25(idx)
    d_ref = single.search_dense(q, k)
    f_ref = single.search_hybrid(q, terms, k, method="rrf", rrf_k=60)
    c_ref = single.search_hybrid(q, terms, k, method="comb_sum", rrf_k=60, w_dense=0.7, w_sparse=0.3)
    for rank in range(2):
        r = np.load(os.path.join(tmp_path, f"rank{rank}.npz"))
        assert np.array_equal(r["d_ids"], d_ref[0]) and np.array_equal(r["d_sc"], d_ref[1])
        assert np.array_equal(r["f_ids"], f_ref[0]) and np.array_equal(r["f_sc"], f_ref[1])
        assert np.array_equal(r["c_ids"], c_ref[0]) and np.array_equal(r["c_sc"], c_ref[1])
        assert np.array_equal(r["f_cnt"], f_ref[3])
        assert np.array_equal(r["g_ids"], f_ref[0]) and np.array_equal(r["g_sc"], f_ref[1])   # sharded build == global build
        assert np.array_equal(r["g_cnt"], f_ref[3])


def _worker_2d(rank, world, port, tmpdir):
    """C = 2 corpus shards x 2 query groups on 4 ranks: every group runs its own all-gather on its own communicator."""
    sys.path.insert(0, os.path.dirname(HERE))
    sys.path.insert(0, HERE)
    import torch.distributed as dist

    from oracle_engine import OracleEngineTorch
    from sentio_b200 import synth
    from sentio_b200.index import build_bm25_from_token_ids
    from sentio_b200.pipeline import HybridPipeline, plan_layout

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    C, n_groups = plan_layout(world, 1.0, corpus_shards=2)
    my_group, r_in = rank // C, rank % C
    group = None
    for g in range(n_groups):  # every rank creates every communicator, in the same order
        pg = dist.new_group(list(range(g * C, (g + 1) * C)))
        if g == my_group:
            group = pg
    n, d, k, B = 2503, 32, 15, 6
    x = synth.dense_corpus(n, d)
    flat, off = synth.text_corpus_tokens(n, vocab=300)
    idx = build_bm25_from_token_ids(flat, off)
    q_all = synth.query_vectors(B * n_groups, d)
    t_all = [idx.term_ids(t) for t in synth.query_tokens(B * n_groups, vocab=300)]
    q, terms = q_all[my_group * B:(my_group + 1) * B], t_all[my_group * B:(my_group + 1) * B]
    lo, hi = (n * r_in) // C, (n * (r_in + 1)) // C
    pipe = HybridPipeline(device=None, rank=r_in, world=C, group=group, engine=OracleEngineTorch())
    pipe.load_dense(x[lo:hi], id_base=lo)
    pipe.load_bm25(idx.shard(lo, hi), id_base=lo)
    f_ids, f_sc, _, f_cnt = pipe.search_hybrid(q, terms, k, method="rrf", rrf_k=60)
    np.savez(os.path.join(tmpdir, f"rank2d{rank}.npz"), f_ids=f_ids, f_sc=f_sc, f_cnt=f_cnt, group=my_group)
    dist.barrier()
    dist.destroy_process_group()


def test_corpus_shards_times_query_groups_layout(tmp_path):
    import torch.multiprocessing as mp

    sys.path.insert(0, HERE)
    from oracle_engine import OracleEngineTorch
    from sentio_b200 import synth
    from sentio_b200.index import build_bm25_from_token_ids
    from sentio_b200.pipeline import HybridPipeline

    port = _free_port()
    mp.spawn(_worker_2d, args=(4, port, str(tmp_path)), nprocs=4, join=True)
    n, d, k, B = 2503, 32, 15, 6
    x = synth.dense_corpus(n, d)
    flat, off = synth.text_corpus_tokens(n, vocab=300)
    idx = build_bm25_from_token_ids(flat, off)
    q_all = synth.query_vectors(B * 2, d)
    t_all = [idx.term_ids(t) for t in synth.query_tokens(B * 2, vocab=300)]
    single = HybridPipeline(device=None, engine=OracleEngineTorch())
    single.load_dense(x)
    single.load_bm25(idx)
    ref = single.search_hybrid(q_all, t_all, k, method="rrf", rrf_k=60)
    for rank in range(4):
        r = np.load(os.path.join(tmp_path, f"rank2d{rank}.npz"))
        g = int(r["group"])
        assert g == rank // 2
        assert np.array_equal(r["f_ids"], ref[0][g * B:(g + 1) * B]) and np.array_equal(r["f_sc"], ref[1][g * B:(g + 1) * B])
        assert np.array_equal(r["f_cnt"], ref[3][g * B:(g + 1) * B])
