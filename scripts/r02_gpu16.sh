#!/bin/bash
# r02 run 17 (1 GPU): BIAS_RES16 epilogue with the residual fetched before the accumulator wait and a single staging pass
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/*.log gpurun_out/*.json gpurun_out/*.csv gpurun_out/status*.txt gpurun_out/*.err
timeout 600 python -m pytest tests/test_rerank_gpu.py tests/test_embedder_gpu.py -m gpu -x -q --timeout=600 > gpurun_out/pytest_ce.log 2>&1
echo "pytest_ce rc=$?" > gpurun_out/status.txt
for v in a b; do
  timeout 600 python bench.py --workload rerank --cpu-sample 0 --no-extras > gpurun_out/bench_rerank_$v.json 2> gpurun_out/bench_rerank_$v.err
  echo "bench $v rc=$?" >> gpurun_out/status.txt
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv --log-file gpurun_out/launches_rerank.csv python bench.py --workload rerank --steps 1 --warmup 1 --inner 1 --cpu-sample 0 --no-extras > gpurun_out/ncu_launch_rerank.log 2>&1
cat gpurun_out/status.txt; tail -3 gpurun_out/pytest_ce.log | cut -c1-300
for v in a b; do tail -1 gpurun_out/bench_rerank_$v.json | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); r=d['roofline']
    print('$v', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ce frac', r['cross_encoder']['frac'], d['clocks']['sm_mhz'])
except Exception as e: print('$v', 'no json', e)"; done
