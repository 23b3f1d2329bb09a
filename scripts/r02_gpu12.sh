#!/bin/bash
# r02 run 13 (1 GPU): BM25 collect pass with 16 queries per CTA over one doc span (dense-row strips shared through L1)
# vs the one-query-per-CTA mapping; parity tests, hybrid A/B, ncu --set full of both collect kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/*.log gpurun_out/*.json gpurun_out/*.csv gpurun_out/status*.txt gpurun_out/*.err
timeout 900 python -m pytest tests/test_bm25_gpu.py tests/test_bm25_build_gpu.py tests/test_hybrid_e2e.py -m gpu -x -q --timeout=600 > gpurun_out/pytest_bm25.log 2>&1
echo "pytest_bm25 rc=$?" > gpurun_out/status.txt
SB_BM25_LOCKSTEP=0 timeout 900 python -m pytest tests/test_bm25_gpu.py -m gpu -x -q --timeout=600 > gpurun_out/pytest_bm25_nolock.log 2>&1
echo "pytest_bm25_nolock rc=$?" >> gpurun_out/status.txt
for v in g0 g1l0 g1l1; do
  G=1; L=1
  [ $v = g0 ] && G=0
  [ $v = g1l0 ] && L=0
  SB_BM25_GROUP=$G SB_BM25_LOCKSTEP=$L timeout 600 python bench.py --workload hybrid --cpu-sample 0 --no-extras > gpurun_out/ab_bm25_$v.json 2> gpurun_out/ab_bm25_$v.err
  echo "bench $v rc=$?" >> gpurun_out/status.txt
done
timeout 500 ncu --set full --clock-control none --import-source on -k regex:bm25_range_kernel -s 2 -c 2 -o gpurun_out/prof_bm25_grouped python bench.py --workload hybrid --steps 1 --warmup 1 --inner 1 --cpu-sample 0 --no-extras > gpurun_out/ncu_full_bm25_grouped.log 2>&1
echo "ncu grouped rc=$?" >> gpurun_out/status.txt
SB_BM25_GROUP=0 timeout 500 ncu --set full --clock-control none --import-source on -k regex:bm25_range_kernel -s 2 -c 2 -o gpurun_out/prof_bm25_single python bench.py --workload hybrid --steps 1 --warmup 1 --inner 1 --cpu-sample 0 --no-extras > gpurun_out/ncu_full_bm25_single.log 2>&1
echo "ncu single rc=$?" >> gpurun_out/status.txt
cat gpurun_out/status.txt; tail -3 gpurun_out/pytest_bm25.log | cut -c1-300; tail -3 gpurun_out/pytest_bm25_nolock.log | cut -c1-300
for v in g0 g1l0 g1l1; do tail -1 gpurun_out/ab_bm25_$v.json | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); r=d['roofline']
    print('$v', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'bm25', r.get('bm25'), d['clocks']['sm_mhz'])
except Exception as e: print('$v', 'no json', e)"; done
