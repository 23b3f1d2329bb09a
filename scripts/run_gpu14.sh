#!/bin/bash
# run 14 (1 GPU): compute-sanitizer on the failing GPU-build test, lean bm25_range_kernel, selector test
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/*.log gpurun_out/*.json gpurun_out/*.csv gpurun_out/*.txt gpurun_out/*.err
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_bm25_gpu.py -m gpu -q -x -k "retriever_class" > gpurun_out/sanitizer.log 2>&1
echo "sanitizer rc=$?" > gpurun_out/status.txt
timeout 900 python -m pytest tests/test_bm25_gpu.py tests/test_bm25_build_gpu.py tests/test_selector_gpu.py -m gpu -q --timeout=600 -s > gpurun_out/pytest_bm25.log 2>&1
echo "pytest_bm25 rc=$?" >> gpurun_out/status.txt
timeout 600 python bench.py --workload hybrid --steps 10 --warmup 3 --cpu-sample 0 > gpurun_out/bench_hybrid.json 2> gpurun_out/bench_hybrid.err
echo "bench hybrid rc=$?" >> gpurun_out/status.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file gpurun_out/launches_hybrid.csv python bench.py --workload hybrid --steps 2 --warmup 1 --cpu-sample 0 > gpurun_out/ncu_launch_hybrid.log 2>&1
echo "ncu launches hybrid rc=$?" >> gpurun_out/status.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bm25_range_kernel -s 3 -c 1 -o gpurun_out/prof_bm25_range python bench.py --workload hybrid --steps 1 --warmup 1 --cpu-sample 0 > gpurun_out/ncu_full_bm25.log 2>&1
echo "ncu full bm25 rc=$?" >> gpurun_out/status.txt
grep -B2 -A25 "Invalid\|ERROR SUMMARY" gpurun_out/sanitizer.log | head -60; tail -12 gpurun_out/pytest_bm25.log; cat gpurun_out/status.txt; cut -c1-330 gpurun_out/bench_hybrid.json; echo; tail -3 gpurun_out/bench_hybrid.err
