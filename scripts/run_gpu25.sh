#!/bin/bash
# run 25 (1 GPU): compute-sanitizer memcheck over the medium-corpus parity tests (multi-range BM25 CTAs, GPU index build,
# tcgen05 dense scan, whole-path host entry points)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/*.log gpurun_out/*.txt
timeout 200 compute-sanitizer --tool memcheck --print-limit 8 --error-exitcode 9 python -m pytest \
  tests/test_bm25_gpu.py tests/test_bm25_build_gpu.py tests/test_hybrid_e2e.py tests/test_dense_gpu.py \
  -m gpu -q --timeout=180 \
  -k "medium or many_queries or device_entry or 7000 or without_export or pipeline_batch or 9000" \
  > gpurun_out/sanitizer2.log 2>&1
echo "sanitizer2 rc=$?" > gpurun_out/status.txt
grep -n "ERROR SUMMARY\|Invalid\|passed\|failed\| at .*kernel" gpurun_out/sanitizer2.log | head -20; cat gpurun_out/status.txt
