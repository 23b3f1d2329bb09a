#!/bin/bash
# r02 run 8 (1 GPU): compute-sanitizer memcheck + racecheck over the kernels written / rewritten in round 2
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/sanitizer*.log gpurun_out/status.txt
timeout 900 python -m pytest tests/test_dense_gpu.py -m gpu -x -q --timeout=800 > gpurun_out/pytest_dense.log 2>&1
echo "pytest_dense rc=$?" > gpurun_out/status0.txt; tail -2 gpurun_out/pytest_dense.log | cut -c1-200
timeout 400 compute-sanitizer --tool memcheck --print-limit 8 --error-exitcode 9 python scripts/racecheck_dense.py > gpurun_out/sanitizer_mem_dense.log 2>&1
echo "memcheck dense rc=$?" > gpurun_out/status.txt
timeout 400 compute-sanitizer --tool memcheck --print-limit 8 --error-exitcode 9 python scripts/racecheck_bm25.py > gpurun_out/sanitizer_mem_bm25.log 2>&1
echo "memcheck bm25 rc=$?" >> gpurun_out/status.txt
timeout 400 compute-sanitizer --tool memcheck --print-limit 8 --error-exitcode 9 python -m pytest tests/test_rerank_gpu.py -m gpu -q -x --timeout=300 -k "small_model or control_flow or (gemm and (96 or 700-384 or 513))" > gpurun_out/sanitizer_mem_ce.log 2>&1
echo "memcheck ce rc=$?" >> gpurun_out/status.txt
timeout 500 compute-sanitizer --tool racecheck --print-limit 6 python scripts/racecheck_bm25.py > gpurun_out/sanitizer_race_bm25.log 2>&1
echo "racecheck bm25 rc=$?" >> gpurun_out/status.txt
timeout 500 compute-sanitizer --tool racecheck --print-limit 6 python scripts/racecheck_dense.py > gpurun_out/sanitizer_race_dense.log 2>&1
echo "racecheck dense rc=$?" >> gpurun_out/status.txt
for f in mem_dense mem_bm25 mem_ce race_bm25 race_dense; do echo "== $f"; grep -n "ERROR SUMMARY\|RACECHECK SUMMARY\|Invalid\|hazard\|ok\|passed\|failed\|Error" gpurun_out/sanitizer_$f.log | head -8; done; cat gpurun_out/status0.txt gpurun_out/status.txt
# knob A/B under sustained load (1 s regions): L2 prefetch distance
for cfg in "base:" "nomulti:SB_DENSE_MULTISAMPLE=0" "pf32:SB_DENSE_PREFETCH=32"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env $envs timeout 300 python bench.py --no-extras --cpu-sample 0 --steps 10 > gpurun_out/ab2_$name.json 2> gpurun_out/ab2_$name.err
  tail -1 gpurun_out/ab2_$name.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('$name', round(d['value']), 'e2e', round(d['e2e']['value']), 'scan ms', round(r['avg_launch_ms'],4), 'frac', round(r['frac'],3))"
done
