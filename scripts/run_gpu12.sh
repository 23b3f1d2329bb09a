#!/bin/bash
# run 12 (1 GPU): parallel radix bin selection, 3-pass sample threshold, attention ILP; full pytest; benches; ncu
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/*.log gpurun_out/*.json gpurun_out/*.csv gpurun_out/*.txt gpurun_out/*.err
timeout 900 python -m pytest tests -m gpu -q --timeout=600 -x > gpurun_out/pytest_full.log 2>&1
echo "pytest_full rc=$?" > gpurun_out/status.txt
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_dense.json 2> gpurun_out/bench_dense.err
echo "bench rc=$?" >> gpurun_out/status.txt
timeout 600 python bench.py --workload hybrid --steps 10 --warmup 3 --cpu-sample 0 > gpurun_out/bench_hybrid.json 2> gpurun_out/bench_hybrid.err
echo "bench hybrid rc=$?" >> gpurun_out/status.txt
timeout 900 python bench.py --workload rerank --steps 5 --warmup 3 --cpu-sample 0 > gpurun_out/bench_rerank.json 2> gpurun_out/bench_rerank.err
echo "bench rerank rc=$?" >> gpurun_out/status.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file gpurun_out/launches_hybrid.csv python bench.py --workload hybrid --steps 2 --warmup 1 --cpu-sample 0 > gpurun_out/ncu_launch_hybrid.log 2>&1
echo "ncu launches hybrid rc=$?" >> gpurun_out/status.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 170 --csv --log-file gpurun_out/launches_rerank.csv python bench.py --workload rerank --steps 1 --warmup 1 --batch 16 --cpu-sample 0 > gpurun_out/ncu_launch_rerank.log 2>&1
echo "ncu launches rerank rc=$?" >> gpurun_out/status.txt
tail -5 gpurun_out/pytest_full.log; cat gpurun_out/status.txt; cut -c1-330 gpurun_out/bench_dense.json; echo; cut -c1-330 gpurun_out/bench_hybrid.json; echo; cut -c1-330 gpurun_out/bench_rerank.json; tail -3 gpurun_out/bench_rerank.err
