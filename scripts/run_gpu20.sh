#!/bin/bash
# run 20 (1 GPU): final default-flag bench lines (batch 256 / 64), reference arm, launch list, quick test subset
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/*.log gpurun_out/*.json gpurun_out/*.csv gpurun_out/*.txt gpurun_out/*.err
timeout 600 python -m pytest tests/test_hybrid_e2e.py tests/test_fuse_scorers_gpu.py tests/test_selector_gpu.py tests/test_abi.py -q --timeout=500 > gpurun_out/pytest_sub.log 2>&1
echo "pytest_sub rc=$?" > gpurun_out/status.txt
timeout 400 python bench.py > gpurun_out/bench_dense.json 2> gpurun_out/bench_dense.err
echo "bench rc=$?" >> gpurun_out/status.txt
timeout 600 python bench.py --workload hybrid --steps 10 --warmup 3 --cpu-sample 0 > gpurun_out/bench_hybrid.json 2> gpurun_out/bench_hybrid.err
echo "bench hybrid rc=$?" >> gpurun_out/status.txt
timeout 900 python bench.py --workload rerank --steps 5 --warmup 3 --cpu-sample 0 > gpurun_out/bench_rerank.json 2> gpurun_out/bench_rerank.err
echo "bench rerank rc=$?" >> gpurun_out/status.txt
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
echo "bench reference rc=$?" >> gpurun_out/status.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_dense.csv python bench.py --steps 2 --warmup 1 --cpu-sample 0 > gpurun_out/ncu_launch_dense.log 2>&1
echo "ncu launches dense rc=$?" >> gpurun_out/status.txt
tail -3 gpurun_out/pytest_sub.log | cut -c1-200; cat gpurun_out/status.txt; for f in dense hybrid rerank reference; do tail -1 gpurun_out/bench_$f.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$f', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms', round(d['ms_per_step'],3), d.get('cpu_baseline') and d['cpu_baseline']['value'], d.get('clocks'))"; done
