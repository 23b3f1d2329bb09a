#!/bin/bash
# run 22 (1 GPU): sb_hybrid_topk host entry point -- parity test + hybrid default bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/*.log gpurun_out/*.json gpurun_out/*.txt gpurun_out/*.err
timeout 400 python -m pytest tests/test_hybrid_e2e.py tests/test_rerank_gpu.py -m gpu -q --timeout=300 -k "pipeline or gpu_stack" > gpurun_out/pytest_sub.log 2>&1
echo "pytest_sub rc=$?" > gpurun_out/status.txt
timeout 600 python bench.py --workload hybrid --cpu-sample 0 > gpurun_out/bench_hybrid.json 2> gpurun_out/bench_hybrid.err
echo "bench hybrid rc=$?" >> gpurun_out/status.txt
tail -3 gpurun_out/pytest_sub.log | cut -c1-300; cat gpurun_out/status.txt; tail -2 gpurun_out/bench_hybrid.err; tail -1 gpurun_out/bench_hybrid.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('hybrid', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms', round(d['ms_per_step'],3), d['config']['batch_queries_per_step'])"
