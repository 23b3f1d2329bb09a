#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/*.log gpurun_out/*.json gpurun_out/*.csv gpurun_out/*.txt
timeout 900 python -m pytest tests -m gpu -q --timeout=600 -k "not full_size" > gpurun_out/pytest_small.log 2>&1
echo "pytest_small rc=$?" > gpurun_out/status.txt
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_dense.json 2> gpurun_out/bench_dense.err
echo "bench rc=$?" >> gpurun_out/status.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --cpu-sample 1 > gpurun_out/ncu_launch.log 2>&1
echo "ncu launches rc=$?" >> gpurun_out/status.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:dense_scan -s 4 -c 1 -o gpurun_out/prof_dense_scan3 python bench.py --steps 2 --warmup 1 --cpu-sample 1 > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?" >> gpurun_out/status.txt
timeout 500 python -m pytest tests -m gpu -q --timeout=450 -k "full_size" > gpurun_out/pytest_full.log 2>&1
echo "pytest_full rc=$?" >> gpurun_out/status.txt
tail -30 gpurun_out/pytest_small.log; tail -3 gpurun_out/pytest_full.log; cat gpurun_out/status.txt; cat gpurun_out/bench_dense.json
