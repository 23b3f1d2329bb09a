#!/bin/bash
# run 21 (1 GPU): hybrid default (batch 128) bench line + ncu --set full of the dense scan and the FFN-up GEMM
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/*.log gpurun_out/*.json gpurun_out/*.csv gpurun_out/*.txt gpurun_out/*.err
timeout 600 python bench.py --workload hybrid --cpu-sample 0 > gpurun_out/bench_hybrid.json 2> gpurun_out/bench_hybrid.err
echo "bench hybrid rc=$?" > gpurun_out/status.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dense_scan_mma_kernel -s 5 -c 1 -o gpurun_out/prof_dense_scan_mma python bench.py --steps 2 --warmup 1 --cpu-sample 0 > gpurun_out/ncu_full_scan.log 2>&1
echo "ncu full scan rc=$?" >> gpurun_out/status.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ce_gemm_ws_kernel -s 8 -c 4 -o gpurun_out/prof_ce_gemm python bench.py --workload rerank --steps 1 --warmup 1 --batch 16 --cpu-sample 0 > gpurun_out/ncu_full_gemm.log 2>&1
echo "ncu full gemm rc=$?" >> gpurun_out/status.txt
cat gpurun_out/status.txt; tail -1 gpurun_out/bench_hybrid.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('hybrid', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms', round(d['ms_per_step'],3), d['config']['batch_queries_per_step'])"
