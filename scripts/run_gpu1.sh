#!/bin/bash
# first GPU session: parity tests, smoke, bench, ncu launch list + full capture of the scan kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import os; print('cpus', os.cpu_count())" >> gpurun_out/gpu.txt
timeout 600 python -m pytest tests -m gpu -q -x --timeout=500 -k "not full_size" > gpurun_out/pytest_small.log 2>&1
echo "pytest_small rc=$?" >> gpurun_out/gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/gpu.txt
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_dense.json 2> gpurun_out/bench_dense.err
echo "bench rc=$?" >> gpurun_out/gpu.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --cpu-sample 1 > gpurun_out/ncu_launch.log 2>&1
echo "ncu launches rc=$?" >> gpurun_out/gpu.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:dense_scan -s 4 -c 2 -o gpurun_out/prof_dense_scan python bench.py --steps 2 --warmup 1 --cpu-sample 1 > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?" >> gpurun_out/gpu.txt
timeout 500 python -m pytest tests -m gpu -q --timeout=450 -k "full_size" > gpurun_out/pytest_full.log 2>&1
echo "pytest_full rc=$?" >> gpurun_out/gpu.txt
tail -5 gpurun_out/pytest_small.log; cat gpurun_out/gpu.txt; cat gpurun_out/bench_dense.json
