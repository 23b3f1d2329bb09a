#!/bin/bash
# r02 run 9 (1 GPU): cross-encoder GEMM with 4 accumulator stages + register bias, BM25 predicated update: parity + bench,
# racecheck of the BM25 kernel again
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/*.log gpurun_out/status*.txt
timeout 1200 python -m pytest tests/test_rerank_gpu.py tests/test_bm25_gpu.py tests/test_hybrid_e2e.py tests/test_embedder_gpu.py tests/test_reference_nodes.py -m gpu -x -q --timeout=900 > gpurun_out/pytest_a.log 2>&1
echo "pytest_a rc=$?" > gpurun_out/status.txt
timeout 600 python bench.py --workload rerank --no-extras --cpu-sample 0 > gpurun_out/bench_rerank.json 2> gpurun_out/bench_rerank.err
echo "bench rerank rc=$?" >> gpurun_out/status.txt
timeout 600 python bench.py --workload hybrid --no-extras --cpu-sample 0 > gpurun_out/bench_hybrid.json 2> gpurun_out/bench_hybrid.err
echo "bench hybrid rc=$?" >> gpurun_out/status.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv --log-file gpurun_out/launches_rerank.csv python bench.py --workload rerank --steps 1 --warmup 1 --inner 1 --cpu-sample 0 --no-extras > gpurun_out/ncu_launch_rerank.log 2>&1
timeout 500 compute-sanitizer --tool racecheck --print-limit 6 python scripts/racecheck_bm25.py > gpurun_out/sanitizer_race_bm25.log 2>&1
echo "racecheck bm25 rc=$?" >> gpurun_out/status.txt
tail -3 gpurun_out/pytest_a.log | cut -c1-300; cat gpurun_out/status.txt; grep -n "RACECHECK SUMMARY\|Race reported\|ok$" gpurun_out/sanitizer_race_bm25.log | head -8 | cut -c1-250
for f in bench_rerank bench_hybrid; do tail -1 gpurun_out/$f.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d.get('roofline') or {}
print('$f', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms/step', round(d['ms_per_step'],3))
for k in ('bm25','cross_encoder'):
    if r.get(k): print('   ', k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in r[k].items() if a in ('postings_per_s','ms_total','share_of_step','achieved','frac')})"; done
