#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/*.log gpurun_out/*.json gpurun_out/*.csv
timeout 900 python -m pytest tests -m gpu -q --timeout=600 -k "not full_size" > gpurun_out/pytest_small.log 2>&1
echo "pytest_small rc=$?" > gpurun_out/status.txt
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_dense.json 2> gpurun_out/bench_dense.err
echo "bench rc=$?" >> gpurun_out/status.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --cpu-sample 1 > gpurun_out/ncu_launch.log 2>&1
echo "ncu launches rc=$?" >> gpurun_out/status.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:dense_scan -s 4 -c 1 -o gpurun_out/prof_dense_scan2 python bench.py --steps 2 --warmup 1 --cpu-sample 1 > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?" >> gpurun_out/status.txt
timeout 600 python bench.py --workload hybrid --steps 10 --warmup 3 --cpu-sample 3 > gpurun_out/bench_hybrid.json 2> gpurun_out/bench_hybrid.err
echo "bench hybrid rc=$?" >> gpurun_out/status.txt
tail -30 gpurun_out/pytest_small.log; cat gpurun_out/status.txt; cat gpurun_out/bench_dense.json; cat gpurun_out/bench_hybrid.json; tail -5 gpurun_out/bench_hybrid.err
