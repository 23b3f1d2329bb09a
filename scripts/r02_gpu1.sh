#!/bin/bash
# r02 run 1 (1 GPU): pair-kernel sanity, dense parity tests, dense bench (pair on / off), launch list
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/*.log gpurun_out/*.json gpurun_out/*.csv gpurun_out/*.txt gpurun_out/*.err
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt
timeout 240 python scripts/r02_pair_check.py > gpurun_out/pair_check.log 2>&1
echo "pair_check rc=$?" > gpurun_out/status.txt
timeout 1500 python -m pytest tests/test_dense_gpu.py -x -q --timeout=900 > gpurun_out/pytest_dense.log 2>&1
echo "pytest_dense rc=$?" >> gpurun_out/status.txt
timeout 400 python bench.py --cpu-sample 0 > gpurun_out/bench_dense.json 2> gpurun_out/bench_dense.err
echo "bench rc=$?" >> gpurun_out/status.txt
SB_DENSE_PAIR=0 timeout 400 python bench.py --cpu-sample 0 > gpurun_out/bench_dense_nopair.json 2> gpurun_out/bench_dense_nopair.err
echo "bench nopair rc=$?" >> gpurun_out/status.txt
SB_DENSE_SAMPLE=1 timeout 400 python bench.py --cpu-sample 0 > gpurun_out/bench_dense_s1.json 2> gpurun_out/bench_dense_s1.err
SB_DENSE_SAMPLE=4 timeout 400 python bench.py --cpu-sample 0 > gpurun_out/bench_dense_s4.json 2> gpurun_out/bench_dense_s4.err
timeout 400 python bench.py --cpu-sample 0 --batch 512 > gpurun_out/bench_dense_b512.json 2> gpurun_out/bench_dense_b512.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_dense.csv python bench.py --steps 2 --warmup 1 --cpu-sample 0 > gpurun_out/ncu_launch_dense.log 2>&1
echo "ncu launches dense rc=$?" >> gpurun_out/status.txt
cat gpurun_out/pair_check.log | tail -5; tail -5 gpurun_out/pytest_dense.log | cut -c1-300; cat gpurun_out/status.txt
for f in dense dense_nopair dense_s1 dense_s4 dense_b512; do tail -1 gpurun_out/bench_$f.json | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read())
    print('$f', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms', round(d['ms_per_step'],3), d.get('roofline'))
except Exception as e: print('$f', 'no json', e)"; done
