#!/bin/bash
# run 23 (1 GPU): sb_hybrid_rerank_topk host entry point -- parity test + rerank default bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/*.log gpurun_out/*.json gpurun_out/*.txt gpurun_out/*.err
timeout 400 python -m pytest tests/test_hybrid_e2e.py tests/test_rerank_gpu.py -m gpu -q --timeout=300 -k "pipeline or gpu_stack" > gpurun_out/pytest_sub.log 2>&1
echo "pytest_sub rc=$?" > gpurun_out/status.txt
timeout 600 python bench.py --workload rerank --steps 5 --warmup 3 --cpu-sample 0 > gpurun_out/bench_rerank.json 2> gpurun_out/bench_rerank.err
echo "bench rerank rc=$?" >> gpurun_out/status.txt
tail -3 gpurun_out/pytest_sub.log | cut -c1-300; cat gpurun_out/status.txt; tail -2 gpurun_out/bench_rerank.err; tail -1 gpurun_out/bench_rerank.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('rerank', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms', round(d['ms_per_step'],3), d['roofline']['cross_encoder']['frac'])"
