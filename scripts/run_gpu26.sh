#!/bin/bash
# run 26 (1 GPU): compute-sanitizer racecheck (shared-memory hazards) on the BM25 range kernel and the attention kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/*.log gpurun_out/*.txt
timeout 110 compute-sanitizer --tool racecheck --print-limit 6 python scripts/racecheck_bm25.py > gpurun_out/racecheck_bm25.log 2>&1
echo "racecheck bm25 rc=$?" > gpurun_out/status.txt
timeout 80 compute-sanitizer --tool racecheck --print-limit 6 python -m pytest tests/test_rerank_gpu.py -m gpu -q -k "small_model" > gpurun_out/racecheck_ce.log 2>&1
echo "racecheck ce rc=$?" >> gpurun_out/status.txt
grep -n "RACECHECK SUMMARY\|hazard\|ok\|passed\|failed\|Error" gpurun_out/racecheck_bm25.log | head -12; grep -n "RACECHECK SUMMARY\|hazard\|passed\|failed" gpurun_out/racecheck_ce.log | head -8; cat gpurun_out/status.txt
