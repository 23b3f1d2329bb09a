#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/*.log gpurun_out/*.json gpurun_out/*.csv gpurun_out/*.txt gpurun_out/*.err
timeout 900 python -m pytest tests -m gpu -q --timeout=600 -k "not full_size" > gpurun_out/pytest_small.log 2>&1
echo "pytest_small rc=$?" > gpurun_out/status.txt
timeout 900 python bench.py --workload rerank --steps 5 --warmup 3 --cpu-sample 0 > gpurun_out/bench_rerank.json 2> gpurun_out/bench_rerank.err
echo "bench rerank rc=$?" >> gpurun_out/status.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 170 --csv --log-file gpurun_out/launches_rerank.csv python bench.py --workload rerank --steps 1 --warmup 1 --batch 16 --cpu-sample 0 > gpurun_out/ncu_launch_rerank.log 2>&1
echo "ncu launches rerank rc=$?" >> gpurun_out/status.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:ce_gemm_ws -s 20 -c 3 -o gpurun_out/prof_ce_gemm_ws python bench.py --workload rerank --steps 1 --warmup 1 --batch 16 --cpu-sample 0 > gpurun_out/ncu_full_gemm.log 2>&1
echo "ncu full gemm rc=$?" >> gpurun_out/status.txt
tail -15 gpurun_out/pytest_small.log; cat gpurun_out/status.txt; cat gpurun_out/bench_rerank.json
