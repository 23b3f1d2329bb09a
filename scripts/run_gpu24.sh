#!/bin/bash
# run 24 (1 GPU): compute-sanitizer memcheck over small-corpus parity tests of every kernel family
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/*.log gpurun_out/*.txt
timeout 330 compute-sanitizer --tool memcheck --print-limit 8 --error-exitcode 9 python -m pytest \
  tests/test_bm25_gpu.py tests/test_bm25_build_gpu.py tests/test_selector_gpu.py tests/test_dense_gpu.py tests/test_rerank_gpu.py tests/test_fuse_scorers_gpu.py \
  -m gpu -q -x --timeout=300 \
  -k "golden or negative_idf or rejects or 33-12 or select_dev or ties or larger or small_model or control_flow or fus" \
  > gpurun_out/sanitizer.log 2>&1
echo "sanitizer rc=$?" > gpurun_out/status.txt
grep -n "ERROR SUMMARY\|Invalid\|passed\|failed\| at .*kernel" gpurun_out/sanitizer.log | head -20; cat gpurun_out/status.txt
