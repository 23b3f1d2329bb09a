#!/bin/bash
# r02 run 20 (1 GPU): ncu --set full of the dominant kernel (pair scan) on the final build -> roofline.traffic
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/*.log gpurun_out/status*.txt
timeout 240 ncu --set full --clock-control none --import-source on -k regex:dense_scan_mma2 -s 4 -c 1 -o gpurun_out/prof_dense_scan_mma2_final python bench.py --steps 1 --warmup 1 --inner 2 --cpu-sample 0 --no-extras > gpurun_out/ncu_full_scan.log 2>&1
echo "ncu full scan rc=$?" > gpurun_out/status.txt
cat gpurun_out/status.txt; tail -3 gpurun_out/ncu_full_scan.log | cut -c1-200
