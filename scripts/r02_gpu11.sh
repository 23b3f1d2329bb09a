#!/bin/bash
# r02 run 11 (1 GPU): experiment -- software completion forwarding in the pair scan (SB_DENSE_SWFWD=1) vs cta_group::2 TMA
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/ab3_* gpurun_out/pytest_swfwd.log
SB_DENSE_SWFWD=1 timeout 900 python -m pytest tests/test_dense_gpu.py -m gpu -x -q --timeout=600 -k "batched or near_tie or fallback or scale" > gpurun_out/pytest_swfwd.log 2>&1
echo "pytest swfwd rc=$?"; tail -2 gpurun_out/pytest_swfwd.log | cut -c1-200
for cfg in "base:" "swfwd:SB_DENSE_SWFWD=1"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env $envs timeout 300 python bench.py --no-extras --cpu-sample 0 --steps 10 > gpurun_out/ab3_$name.json 2> gpurun_out/ab3_$name.err
  tail -1 gpurun_out/ab3_$name.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('$name', round(d['value']), 'e2e', round(d['e2e']['value']), 'scan ms', round(r['avg_launch_ms'],4), 'frac', round(r['frac'],3))"
done
