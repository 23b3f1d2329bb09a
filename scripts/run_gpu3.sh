#!/bin/bash
# GPU session: parity tests, bench (dense, batch 64 -> tcgen05 scan; batch 4 -> CUDA-core scan), ncu launch list + captures
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/*.log gpurun_out/*.json gpurun_out/*.csv gpurun_out/*.txt
timeout 900 python -m pytest tests -m gpu -q --timeout=600 -k "not full_size" > gpurun_out/pytest_small.log 2>&1
echo "pytest_small rc=$?" > gpurun_out/status.txt
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_dense.json 2> gpurun_out/bench_dense.err
echo "bench rc=$?" >> gpurun_out/status.txt
timeout 400 python bench.py --steps 20 --warmup 3 --batch 4 --cpu-sample 1 > gpurun_out/bench_dense_b4.json 2> gpurun_out/bench_dense_b4.err
echo "bench b4 rc=$?" >> gpurun_out/status.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --cpu-sample 1 > gpurun_out/ncu_launch.log 2>&1
echo "ncu launches rc=$?" >> gpurun_out/status.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:dense_scan -s 6 -c 1 -o gpurun_out/prof_dense_scan_mma python bench.py --steps 2 --warmup 1 --cpu-sample 1 > gpurun_out/ncu_full.log 2>&1
echo "ncu full mma rc=$?" >> gpurun_out/status.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:dense_scan -s 4 -c 1 -o gpurun_out/prof_dense_scan_b4 python bench.py --steps 2 --warmup 1 --batch 4 --cpu-sample 1 > gpurun_out/ncu_full_b4.log 2>&1
echo "ncu full b4 rc=$?" >> gpurun_out/status.txt
timeout 500 python -m pytest tests -m gpu -q --timeout=450 -k "full_size" > gpurun_out/pytest_full.log 2>&1
echo "pytest_full rc=$?" >> gpurun_out/status.txt
tail -30 gpurun_out/pytest_small.log; tail -3 gpurun_out/pytest_full.log; cat gpurun_out/status.txt; cat gpurun_out/bench_dense.json; cat gpurun_out/bench_dense_b4.json
