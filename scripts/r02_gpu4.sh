#!/bin/bash
# r02 run 4 (1 GPU): parity tests after the fixes, ncu --set full of the three hot kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/*.log gpurun_out/*.json gpurun_out/*.csv gpurun_out/*.txt gpurun_out/*.err gpurun_out/*.ncu-rep
timeout 1200 python -m pytest tests/test_rerank_gpu.py tests/test_reference_nodes.py tests/test_embedder_gpu.py tests/test_selector_gpu.py tests/test_fuse_scorers_gpu.py tests/test_abi.py -m gpu -x -q --timeout=900 > gpurun_out/pytest_a.log 2>&1
echo "pytest_a rc=$?" > gpurun_out/status.txt
timeout 900 python -m pytest tests/test_dense_gpu.py -m gpu -x -q --timeout=900 -k "not full_size" > gpurun_out/pytest_dense.log 2>&1
echo "pytest_dense rc=$?" >> gpurun_out/status.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:dense_scan_mma2 -s 2 -c 1 -o gpurun_out/prof_dense_scan_mma2 python bench.py --steps 1 --warmup 1 --inner 2 --cpu-sample 0 --no-extras > gpurun_out/ncu_full_scan.log 2>&1
echo "ncu full scan rc=$?" >> gpurun_out/status.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:dense_select_kernel -s 1 -c 1 -o gpurun_out/prof_dense_select python bench.py --steps 1 --warmup 1 --inner 2 --cpu-sample 0 --no-extras > gpurun_out/ncu_full_select.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:bm25_range_kernel -s 1 -c 1 -o gpurun_out/prof_bm25_range python bench.py --workload hybrid --steps 1 --warmup 1 --inner 1 --cpu-sample 0 --no-extras > gpurun_out/ncu_full_bm25.log 2>&1
echo "ncu full bm25 rc=$?" >> gpurun_out/status.txt
timeout 500 ncu --set full --clock-control none --import-source on -k "regex:ce_gemm_ws_kernel|ce_attention_mma|ce_ln_kernel" -s 12 -c 7 -o gpurun_out/prof_ce python bench.py --workload rerank --steps 1 --warmup 1 --inner 1 --batch 16 --cpu-sample 0 --no-extras > gpurun_out/ncu_full_ce.log 2>&1
echo "ncu full ce rc=$?" >> gpurun_out/status.txt
timeout 600 python bench.py --workload rerank --no-extras --cpu-sample 0 > gpurun_out/bench_rerank.json 2> gpurun_out/bench_rerank.err
echo "bench rerank rc=$?" >> gpurun_out/status.txt
tail -3 gpurun_out/pytest_a.log | cut -c1-300; tail -3 gpurun_out/pytest_dense.log | cut -c1-300; cat gpurun_out/status.txt; ls -la gpurun_out/*.ncu-rep
tail -1 gpurun_out/bench_rerank.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('rerank', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), r['cross_encoder']['frac'], r['cross_encoder']['achieved'])"
