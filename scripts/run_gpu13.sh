#!/bin/bash
# run 13 (1 GPU): range-ahead prefetch in bm25_range_kernel, GPU index build, head kernel; tests + benches + ncu
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/*.log gpurun_out/*.json gpurun_out/*.csv gpurun_out/*.txt gpurun_out/*.err
timeout 900 python -m pytest tests/test_bm25_gpu.py tests/test_bm25_build_gpu.py -m gpu -q --timeout=600 -x -s > gpurun_out/pytest_bm25.log 2>&1
echo "pytest_bm25 rc=$?" > gpurun_out/status.txt
timeout 900 python -m pytest tests -m gpu -q --timeout=600 -x -k "not full_size" --deselect tests/test_bm25_gpu.py --deselect tests/test_bm25_build_gpu.py > gpurun_out/pytest_rest.log 2>&1
echo "pytest_rest rc=$?" >> gpurun_out/status.txt
timeout 600 python bench.py --workload hybrid --steps 10 --warmup 3 --cpu-sample 0 > gpurun_out/bench_hybrid.json 2> gpurun_out/bench_hybrid.err
echo "bench hybrid rc=$?" >> gpurun_out/status.txt
timeout 900 python bench.py --workload rerank --steps 5 --warmup 3 --cpu-sample 0 > gpurun_out/bench_rerank.json 2> gpurun_out/bench_rerank.err
echo "bench rerank rc=$?" >> gpurun_out/status.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file gpurun_out/launches_hybrid.csv python bench.py --workload hybrid --steps 2 --warmup 1 --cpu-sample 0 > gpurun_out/ncu_launch_hybrid.log 2>&1
echo "ncu launches hybrid rc=$?" >> gpurun_out/status.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bm25_range_kernel -s 3 -c 1 -o gpurun_out/prof_bm25_range python bench.py --workload hybrid --steps 1 --warmup 1 --cpu-sample 0 > gpurun_out/ncu_full_bm25.log 2>&1
echo "ncu full bm25 rc=$?" >> gpurun_out/status.txt
tail -6 gpurun_out/pytest_bm25.log; tail -4 gpurun_out/pytest_rest.log; cat gpurun_out/status.txt; cut -c1-330 gpurun_out/bench_hybrid.json; echo; tail -3 gpurun_out/bench_hybrid.err; cut -c1-330 gpurun_out/bench_rerank.json; tail -3 gpurun_out/bench_rerank.err
