#!/bin/bash
# r02 run 5 (1 GPU): BM25 dense head-term rows + lean walk, cross-encoder transposed pair GEMM: parity, bench, launch lists
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/*.log gpurun_out/*.json gpurun_out/*.csv gpurun_out/*.txt gpurun_out/*.err gpurun_out/*.ncu-rep
timeout 1500 python -m pytest tests/test_rerank_gpu.py tests/test_bm25_gpu.py tests/test_bm25_build_gpu.py tests/test_hybrid_e2e.py tests/test_embedder_gpu.py -m gpu -x -q --timeout=900 > gpurun_out/pytest_a.log 2>&1
echo "pytest_a rc=$?" > gpurun_out/status.txt
timeout 900 python bench.py --workload hybrid --no-extras --cpu-sample 0 > gpurun_out/bench_hybrid.json 2> gpurun_out/bench_hybrid.err
echo "bench hybrid rc=$?" >> gpurun_out/status.txt
SB_BM25_DENSE=0 timeout 900 python bench.py --workload hybrid --no-extras --cpu-sample 0 > gpurun_out/bench_hybrid_nodense.json 2> gpurun_out/bench_hybrid_nodense.err
timeout 900 python bench.py --workload rerank --no-extras --cpu-sample 0 > gpurun_out/bench_rerank.json 2> gpurun_out/bench_rerank.err
echo "bench rerank rc=$?" >> gpurun_out/status.txt
SB_CE_PAIR_GEMM=0 timeout 900 python bench.py --workload rerank --no-extras --cpu-sample 0 > gpurun_out/bench_rerank_nopair.json 2> gpurun_out/bench_rerank_nopair.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv --log-file gpurun_out/launches_rerank.csv python bench.py --workload rerank --steps 1 --warmup 1 --inner 1 --cpu-sample 0 --no-extras > gpurun_out/ncu_launch_rerank.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_hybrid.csv python bench.py --workload hybrid --steps 1 --warmup 1 --inner 1 --cpu-sample 0 --no-extras > gpurun_out/ncu_launch_hybrid.log 2>&1
echo "ncu rc=$?" >> gpurun_out/status.txt
tail -4 gpurun_out/pytest_a.log | cut -c1-300; cat gpurun_out/status.txt
for f in bench_hybrid bench_hybrid_nodense bench_rerank bench_rerank_nopair; do tail -1 gpurun_out/$f.json | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read())
    r=d.get('roofline') or {}
    print('$f', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms/step', round(d['ms_per_step'],3))
    for k in ('bm25','cross_encoder'):
        if r.get(k): print('   ', k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in r[k].items() if a in ('postings_per_s','posting_GBps','ms_total','share_of_step','achieved','frac')})
except Exception as e: print('$f', 'no json', e)"; done
