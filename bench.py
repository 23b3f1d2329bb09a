#!/usr/bin/env python
"""bench.py -- retrieval queries/sec on BASELINE.json's configurations.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--workload dense|hybrid|rerank] [--batch B] [--n-docs N] [--dim D] [--top-k K] [--rerank-k K2]

A "step" = one batch of B synthetic queries through the hot path.  Default workload = BASELINE.json configs[1]:
1 M docs x 1024-d, dense-only cosine top_k=100 on 1 x B200.  --batch is the number of queries per step PER GPU (default 256;
128 for hybrid, 64 for the rerank workload, i.e. 6400 pairs per step): a step on N GPUs carries N x batch queries.  Multi-GPU layout =
C corpus shards x N/C query groups: the C ranks of a group partition the corpus (contiguous doc ranges), score the group's
C x batch queries against their shards, exchange per-shard top-k in ONE NCCL all-gather and merge; different groups answer
different queries.  --shard corpus: C = N (north_star's layout for corpora that must be partitioned); --shard queries:
C = 1 (replicated corpus, no collective); --shard auto (default): the smallest C whose shard fits the per-GPU memory
budget -- 1 for the 2 GB corpus of the metric.  Per-GPU work per step is constant in N -> "scaling": "weak"; the corpus
of the metric (1 M docs) never changes.

value   : whole-job queries/sec, inputs already resident in HBM (device entry points, CUDA-event timed, max over ranks)
e2e     : the same metric through the host-buffer C-ABI entry point (pinned host queries -> H2D -> kernels -> D2H results)
roofline: dominant kernel (dense_scan_kernel) algorithmic bytes / CUDA-event duration vs MEASURED_PEAKS.json hbm_gbs
cpu_baseline / --impl reference: the reference's CPU path (exact cosine in NumPy: fp32 `X @ q` with BLAS on all host
          cores + full np.argsort, oracle/dense.py:fast_topk_f32; BM25 via the rank_bm25 restatement) on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "retrieval queries/sec @1M docs,1024-d,top_k=100"  # BASELINE.json metric (the workload actually run is in config)
UNIT = "queries/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="dense", choices=["dense", "hybrid", "rerank"])
    ap.add_argument("--rerank-k", type=int, default=10, help="documents kept after the cross-encoder (config 4: 100 -> 10)")
    ap.add_argument("--batch", type=int, default=None,
                    help="queries per step PER GPU (default: 256 dense, 128 hybrid, 64 rerank = 6400 pairs per step)")
    ap.add_argument("--n-docs", type=int, default=1_000_000)
    ap.add_argument("--dim", type=int, default=1024)
    ap.add_argument("--top-k", type=int, default=100)
    ap.add_argument("--cpu-sample", type=int, default=24, help="queries in the bounded CPU-baseline sample")
    ap.add_argument("--shard", default="auto", choices=["auto", "corpus", "queries"],
                    help="--gpus N > 1 layout = C corpus shards x N/C query groups.  'corpus': C = N (contiguous doc ranges "
                         "+ ONE NCCL all-gather of per-shard top-k, north_star's layout for corpora that must be "
                         "partitioned); 'queries': C = 1 (corpus replicated, queries split, no collective); 'auto' "
                         "(default): the smallest C whose shard fits --gpu-mem-budget-gb, i.e. partition only as much "
                         "as capacity requires")
    ap.add_argument("--corpus-shards", type=int, default=0, help="explicit C (must divide N); overrides --shard")
    ap.add_argument("--gpu-mem-budget-gb", type=float, default=64.0,
                    help="HBM one GPU may spend on index data under --shard auto (B200: 180 GB)")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.device), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------- synthetic workload
def make_workload(args, lo=None, hi=None):
    """Synthetic corpus / queries.  Corpora above 2 M docs are generated shard-locally from
    independently seeded chunks (synth.dense_corpus_range), so a rank only ever materialises its own rows."""
    from sentio_b200 import synth

    t0 = time.time()
    if args.n_docs > 2_000_000 and lo is not None:
        x16 = synth.dense_corpus_range(lo, hi, args.dim)
    else:
        x16 = synth.dense_corpus(args.n_docs, args.dim)
        if lo is not None:
            x16 = x16[lo:hi]
    queries = synth.query_vectors(1024, args.dim)
    wl = {"x16": x16, "q": queries, "gen_s": None}
    if args.workload in ("hybrid", "rerank"):
        flat, off = synth.text_corpus_tokens(args.n_docs)
        wl["flat"], wl["off"] = flat, off
        wl["q_tokens"] = synth.query_tokens(1024)
    wl["gen_s"] = round(time.time() - t0, 1)
    return wl


# --------------------------------------------------------------------------------------------- CPU reference arm
def cpu_reference(args, wl, n_queries):
    """Times the reference's CPU path on this box's host cores on a bounded sample of the same workload."""
    from oracle import dense as dense_oracle

    cores = os.cpu_count() or 1
    x16 = wl["x16"]
    x32 = x16.astype(np.float32)
    x32 /= np.linalg.norm(x32, axis=1, keepdims=True)  # Qdrant normalises at upsert; the scan is then a plain dot
    q = wl["q"]
    fast = None
    ce_model = None
    if args.workload in ("hybrid", "rerank"):
        from oracle import fusion as fusion_oracle
        from oracle.rank_bm25_port import FastBM25
        from sentio_b200.index import build_bm25_from_token_ids

        idx = build_bm25_from_token_ids(wl["flat"], wl["off"])
        fast = FastBM25(idx.indptr, idx.post_doc, idx.post_tf, idx.doc_len, idx.idf, idx.avgdl)
        terms = [idx.term_ids(t) for t in wl["q_tokens"]]
    if args.workload == "rerank":
        from oracle import cross_encoder as ce_oracle
        from sentio_b200.cross_encoder import MINILM_L6
        from sentio_b200.index import hash_tokenize_pairs
        from sentio_b200 import synth

        ce_model = ce_oracle.hf_model(MINILM_L6, seed=0)
    for i in range(2):  # warm-up
        dense_oracle.fast_topk_f32(x32, q[i], args.top_k)
    t0 = time.perf_counter()
    for i in range(n_queries):
        di, ds = dense_oracle.fast_topk_f32(x32, q[i % len(q)], args.top_k)
        if fast is not None:
            s = fast.get_scores(list(terms[i % len(terms)]))
            order = np.argsort(-s)[: args.top_k]
            sp = [(int(j), float(s[j])) for j in order if s[j] > 0]
            fused = fusion_oracle.fuse("rrf", 60, 0.5, 0.5, [(int(a), float(b)) for a, b in zip(di, ds)], sp, [],
                                       args.top_k)
            if ce_model is not None:
                qtext = synth.token_text(wl["q_tokens"][i % len(terms)])
                texts = [synth.token_text(wl["flat"][wl["off"][d]:wl["off"][d + 1]]) for d, _, _ in fused]
                ids, tt, lens = hash_tokenize_pairs(qtext, texts, 128)
                _, sig = ce_oracle.hf_scores(ce_model, ids, tt, lens, batch=len(texts))
                sorted(range(len(sig)), key=lambda j: -sig[j])[: args.rerank_k]
    dt = time.perf_counter() - t0
    kind = "port"
    sample = (f"{n_queries} queries of the same workload; dense = fp32 X@q (NumPy/BLAS, {cores} threads) + np.argsort"
              + ("; BM25 = CSR restatement of rank_bm25 get_scores + np.argsort; rrf fusion in Python" if fast else "")
              + ("; rerank = HuggingFace BertForSequenceClassification (MiniLM-L6 shape) fp32 on CPU, 100 pairs/query"
                 if ce_model is not None else ""))
    return {"value": n_queries / dt, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample}, dt / n_queries


# --------------------------------------------------------------------------------------------- main
def main():
    args = parse_args()
    if args.batch is None:
        args.batch = {"dense": 256, "hybrid": 128, "rerank": 64}[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    workload_name = (f"{args.n_docs}-doc synthetic, {args.dim}-d, "
                     + {"dense": "dense-only cosine", "hybrid": "hybrid dense+BM25 rrf",
                        "rerank": "hybrid dense+BM25 rrf + cross-encoder rerank (MiniLM-L6 random-init)"}[args.workload]
                     + f" top_k={args.top_k}" + (f"->{args.rerank_k}" if args.workload == "rerank" else ""))
    # ---- multi-GPU layout: C corpus shards x (world / C) query groups
    est_gb = args.n_docs * args.dim * 2 / 1e9 + (args.n_docs * 60 * 12 / 1e9 if args.workload != "dense" else 0.0)
    from sentio_b200.pipeline import plan_layout

    try:
        C, _ = plan_layout(world, est_gb, args.shard, args.corpus_shards, args.gpu_mem_budget_gb)
    except ValueError as exc:
        raise SystemExit(str(exc))
    n_groups, my_group, r_in = world // C, rank // C, rank % C
    sharded = C > 1
    replicated = n_groups > 1
    B_gpu = args.batch
    B_total = args.batch * world
    config = {"workload": workload_name, "batch_queries_per_step": B_total, "queries_per_gpu_per_step": B_gpu,
              "store_dtype": "fp16", "shards": C, "query_groups": n_groups, "index_gb_estimate": round(est_gb, 2),
              "multi_gpu": ("single GPU" if world == 1 else
                            f"{C} corpus shard(s) x {n_groups} query group(s): "
                            + ("corpus partition + one NCCL all-gather of per-shard top-k" if sharded else
                               "corpus replicated, no collective")
                            + ("; queries split across groups" if replicated else "")
                            + (f" [--corpus-shards {C}]" if args.corpus_shards else f" [--shard {args.shard}]")),
              "l2_policy": "corpus (2.05 GB) is larger than L2 (126 MB); no flush needed",
              "query_set": "1024 seeded unit vectors, cycled"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        wl = make_workload(args)
        per_step = max(1, min(args.batch, 4))
        total = per_step * (args.steps + args.warmup)
        base, per_q = cpu_reference(args, wl, total)
        line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_q * per_step * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": {**config, "batch_queries_per_step": per_step},
                "cpu_baseline": {**base, "sample": f"{per_step} queries per step; " + base["sample"]},
                "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist

    from sentio_b200.pipeline import HybridPipeline

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    n = args.n_docs
    if sharded:
        lo, hi = (n * r_in) // C, (n * (r_in + 1)) // C
    else:
        lo, hi = 0, n
    wl = make_workload(args, lo, hi)
    group = None
    if sharded and n_groups > 1:  # one NCCL communicator per corpus group (every rank creates all of them, in order)
        for g in range(n_groups):
            pg = dist.new_group(list(range(g * C, (g + 1) * C)))
            if g == my_group:
                group = pg
    pipe = HybridPipeline(local_rank, rank=r_in if sharded else 0, world=C if sharded else 1, group=group)
    pipe.load_dense(wl["x16"], id_base=lo)
    idx = None
    rerank = args.workload == "rerank"
    if args.workload in ("hybrid", "rerank"):
        from sentio_b200.index import build_bm25_from_token_ids

        if sharded:  # corpus-global idf / avgdl: build once on the host, upload this rank's shard
            idx = build_bm25_from_token_ids(wl["flat"], wl["off"])
            pipe.load_bm25(idx.shard(lo, hi), id_base=lo)
        else:        # single shard: the index is built on the device (sb_bm25_build_*), 0.3 s at 1 M docs
            idx = pipe.engine.build_bm25_gpu(wl["flat"], wl["off"], export=True)
    if rerank:
        from sentio_b200 import synth
        from sentio_b200.cross_encoder import MINILM_L6, CrossEncoderWeights
        from sentio_b200.index import doc_token_matrix, hash_vocab_ids

        vocab_ids = hash_vocab_ids(synth.VOCAB)
        doc_tok, doc_len = doc_token_matrix(wl["flat"], wl["off"], vocab_ids, ld=120)
        pipe.load_cross_encoder(CrossEncoderWeights.random(MINILM_L6, seed=0))
        pipe.load_doc_tokens(doc_tok, doc_len, id_base=0)  # replicated on every rank (240 MB at 1 M docs)
        q_tok_all = vocab_ids[wl["q_tokens"]].astype(np.int32)
    eng = pipe.engine
    # queries this rank handles per step: those of its corpus group (all C ranks of a group score the same queries)
    B, k = B_gpu * C, args.top_k
    q_shift = my_group * B
    dev = f"cuda:{local_rank}"
    q_all = torch.from_numpy(wl["q"]).to(dev)
    n_q = q_all.shape[0]
    terms_dev = None
    if idx is not None:
        term_lists = [idx.term_ids(t) for t in wl["q_tokens"]]

    def batch_slice(step):
        s = (step * B_total + q_shift) % n_q
        idxs = [(s + i) % n_q for i in range(B)]
        return idxs

    def dev_inputs(step):
        ids = batch_slice(step)
        qt = q_all[ids].contiguous()
        if idx is None:
            return (qt,)
        flat, off = eng.pack_queries([term_lists[i] for i in ids])
        base = (qt, torch.from_numpy(flat).to(dev), torch.from_numpy(off).to(dev), int(off[-1]),
                int(np.diff(off).max()))
        if not rerank:
            return base
        qtok = torch.from_numpy(q_tok_all[ids]).to(dev)
        qlen = torch.full((B,), q_tok_all.shape[1], dtype=torch.int32, device=dev)
        return base + (qtok, qlen)

    def run_dev(inp):
        if idx is None:
            return pipe.dense_dev(inp[0], k)
        if rerank:
            return pipe.hybrid_rerank_dev(inp[0], inp[1], inp[2], inp[3], inp[4], inp[5], inp[6], k, args.rerank_k, 128,
                                          "rrf", 60, 0.5, 0.5)
        return pipe.hybrid_dev(inp[0], inp[1], inp[2], inp[3], inp[4], k, "rrf", 60, 0.5, 0.5)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- resident leg: inputs already in HBM, device entry points, CUDA events on torch's current stream
    inputs = [dev_inputs(s) for s in range(args.warmup + args.steps)]
    for s in range(args.warmup):
        run_dev(inputs[s])
    barrier()
    eng.profile(True)
    if rerank:
        eng.ce_stats(reset=True)
    launches0 = eng.launch_count()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for s in range(args.steps):
        run_dev(inputs[args.warmup + s])
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    launches = eng.launch_count() - launches0
    n_scan, scan_ms = eng.profile_read("dense_scan")
    n_ce, ce_ms = eng.profile_read("ce")
    n_bm, bm_ms = eng.profile_read("bm25_score") if idx is not None else (0, 0.0)
    ce_pairs, ce_rows, ce_sq = eng.ce_stats() if rerank else (0, 0, 0)
    eng.profile(False)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = B_total * args.steps / (ms_total / 1e3)

    # ---------------- e2e leg: host (pinned by the library) buffers in, host results out, every step
    host_batches = [wl["q"][batch_slice(s)] for s in range(args.warmup + args.steps)]
    host_terms = None
    if idx is not None:
        host_terms = [[term_lists[i] for i in batch_slice(s)] for s in range(args.warmup + args.steps)]

    def run_host(s):
        if idx is None:
            return pipe.search_dense(host_batches[s], k)
        if rerank:
            sl = batch_slice(s)
            return pipe.search_hybrid_rerank(host_batches[s], host_terms[s], q_tok_all[sl],
                                             np.full(B, q_tok_all.shape[1], np.int32), k, args.rerank_k, 128, "rrf", 60,
                                             0.5, 0.5)
        return pipe.search_hybrid(host_batches[s], host_terms[s], k, "rrf", 60, 0.5, 0.5)

    for s in range(args.warmup):
        run_host(s)
    barrier()
    t0 = time.perf_counter()
    for s in range(args.steps):
        out = run_host(args.warmup + s)
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item())
    h2d = B * args.dim * 4 + (0 if idx is None else sum(len(x) for x in host_terms[0]) * 4 + (B + 1) * 4)
    d2h = B * k * 16 + B * 4 + (B * k * 4 if idx is not None else 0)
    if rerank:
        h2d += B * q_tok_all.shape[1] * 4 + B * 4
        d2h = B_gpu * args.rerank_k * 12 + B_gpu * 4  # every rank returns the rows of the queries it reranked
    e2e = {"value": B_total * args.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d * world,
           "d2h_bytes_per_step": d2h * world,
           "timer": "host wall clock around the public host-buffer call (includes H2D, kernels, D2H, sync); "
                    "bytes are summed over ranks"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---------------- roofline of the dominant kernel (dense_scan_kernel), this rank's shard
    peak, peak_src = peaks()
    rows = hi - lo
    n_pad = (rows + 31) // 32 * 32
    d_pad = (args.dim + 7) // 8 * 8
    alg_bytes = n_pad * d_pad * 2 + n_pad * 4
    avg_ms = scan_ms / max(n_scan, 1)
    achieved = alg_bytes / (avg_ms * 1e-3) / 1e9 if n_scan else 0.0
    scan_kernel = "dense_scan_mma_kernel (tcgen05, batched queries)" if B >= 16 else "dense_scan_kernel (FFMA2)"
    roofline = {"bound": "hbm", "kernel": scan_kernel, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": avg_ms, "launches_timed": n_scan,
                "queries_per_launch": (B * args.steps) / max(n_scan, 1), "share_of_step": scan_ms / ms_total}
    prof = os.path.join(ROOT, "profiles", "r01_dense_scan_ncu.json")
    if os.path.exists(prof) and world == 1:
        try:
            roofline["traffic"] = json.load(open(prof)).get("dram_bytes_per_launch")
        except Exception:
            pass

    if idx is not None and n_bm:
        # BM25 range kernel (sample + collect launches): algorithmic bytes = the postings of the query terms, 12 B each
        # (4 B doc + 8 B fp64 ratio); this rank's shard, all queries of the timed steps
        sidx = idx.shard(lo, hi) if sharded else idx
        df = np.diff(sidx.indptr)
        postings = 0
        for s_ in range(args.steps):
            for i in batch_slice(args.warmup + s_):
                t = term_lists[i]
                postings += int(df[t[t >= 0]].sum())
        bm_bytes = postings * 12
        roofline["bm25"] = {"bound": "hbm", "kernel": "bm25_range_kernel (sample + collect)",
                            "achieved": bm_bytes / (bm_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                            "frac": bm_bytes / (bm_ms * 1e-3) / 1e9 / peak, "algorithmic_bytes": bm_bytes,
                            "postings_per_query": postings / (B * args.steps), "ms_total": bm_ms,
                            "share_of_step": bm_ms / ms_total,
                            "note": "posting lists shared by the queries of a batch are served from L2 (ncu: DRAM "
                                    "traffic ~0.25 GB per 64-query launch vs 1.6 GB algorithmic); the kernel is "
                                    "issue/latency bound, see profiles/r01_run15_bm25_range_ncu.md"}
    if rerank and n_ce:
        # second roofline: the cross-encoder forward (tensor pipe).  Flops of the work actually done: the packed-token
        # forward computes sum(len) token rows, not P x 128 (library counters); 2.87 GFLOP per pair only at len = 128.
        flops = 6 * (24 * 384 * 384 * ce_rows + 4 * 384 * ce_sq)
        tf = flops / (ce_ms * 1e-3) / 1e12
        tpeak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("bf16_tflops_sustained", 1429.5) \
            if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 1400.0
        roofline["cross_encoder"] = {"bound": "tensor", "achieved": tf, "peak": tpeak, "unit": "TFLOP/s",
                                     "frac": tf / tpeak, "ms_total": ce_ms, "forward_calls": n_ce,
                                     "pairs": ce_pairs, "mean_pair_len": ce_rows / max(ce_pairs, 1),
                                     "flops_counted": "L*(24*H^2*sum(len) + 4*H*sum(len^2)), padding excluded",
                                     "share_of_step": ce_ms / ms_total}

    # ---------------- bounded CPU baseline on this box's host cores (rank 0, N=1 only)
    cpu = None
    if world == 1 and args.cpu_sample > 0 and args.n_docs <= 2_000_000:
        cpu, _ = cpu_reference(args, wl, args.cpu_sample)

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f16 store / f32 scan / f64 exact re-score",
            "data": "synthetic", "config": config, "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline, "cpu_baseline": cpu, "corpus_gen_s": wl["gen_s"]}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
