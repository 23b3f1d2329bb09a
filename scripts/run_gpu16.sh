#!/bin/bash
# run 16 (1 GPU): full GPU suite + the three benches + launch lists + ncu full of dense_select / attention
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/*.log gpurun_out/*.json gpurun_out/*.csv gpurun_out/*.txt gpurun_out/*.err
timeout 1200 python -m pytest tests -m gpu -q --timeout=900 > gpurun_out/pytest_full.log 2>&1
echo "pytest_full rc=$?" > gpurun_out/status.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/status.txt
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_dense.json 2> gpurun_out/bench_dense.err
echo "bench rc=$?" >> gpurun_out/status.txt
timeout 600 python bench.py --workload hybrid --steps 10 --warmup 3 --cpu-sample 0 > gpurun_out/bench_hybrid.json 2> gpurun_out/bench_hybrid.err
echo "bench hybrid rc=$?" >> gpurun_out/status.txt
timeout 900 python bench.py --workload rerank --steps 5 --warmup 3 --cpu-sample 0 > gpurun_out/bench_rerank.json 2> gpurun_out/bench_rerank.err
echo "bench rerank rc=$?" >> gpurun_out/status.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_dense.csv python bench.py --steps 2 --warmup 1 --cpu-sample 0 > gpurun_out/ncu_launch_dense.log 2>&1
echo "ncu launches dense rc=$?" >> gpurun_out/status.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file gpurun_out/launches_hybrid.csv python bench.py --workload hybrid --steps 2 --warmup 1 --cpu-sample 0 > gpurun_out/ncu_launch_hybrid.log 2>&1
echo "ncu launches hybrid rc=$?" >> gpurun_out/status.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 170 --csv --log-file gpurun_out/launches_rerank.csv python bench.py --workload rerank --steps 1 --warmup 1 --batch 16 --cpu-sample 0 > gpurun_out/ncu_launch_rerank.log 2>&1
echo "ncu launches rerank rc=$?" >> gpurun_out/status.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dense_select_kernel -s 5 -c 1 -o gpurun_out/prof_dense_select python bench.py --steps 2 --warmup 1 --cpu-sample 0 > gpurun_out/ncu_full_select.log 2>&1
echo "ncu full select rc=$?" >> gpurun_out/status.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ce_attention_mma_kernel -s 7 -c 1 -o gpurun_out/prof_ce_attention python bench.py --workload rerank --steps 1 --warmup 1 --batch 16 --cpu-sample 0 > gpurun_out/ncu_full_attn.log 2>&1
echo "ncu full attention rc=$?" >> gpurun_out/status.txt
tail -6 gpurun_out/pytest_full.log | cut -c1-200; cat gpurun_out/smoke.log | tail -2; cat gpurun_out/status.txt; cut -c1-330 gpurun_out/bench_dense.json; echo; cut -c1-330 gpurun_out/bench_hybrid.json; echo; cut -c1-330 gpurun_out/bench_rerank.json; tail -3 gpurun_out/bench_rerank.err
