#!/bin/bash
# r02 run 16 (1 GPU): cross-encoder GEMM epilogue with register bias + 8-byte drain (default) vs the 4-byte drain
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/*.log gpurun_out/*.json gpurun_out/*.csv gpurun_out/status*.txt gpurun_out/*.err
timeout 600 python -m pytest tests/test_rerank_gpu.py tests/test_embedder_gpu.py -m gpu -x -q --timeout=600 > gpurun_out/pytest_ce.log 2>&1
echo "pytest_ce rc=$?" > gpurun_out/status.txt
for v in drain32 base drain32b base2; do
  lib=$PWD/sentio_b200/libsentio_b200.so
  [ ${v:0:7} = drain32 ] && lib=$PWD/sentio_b200/libsentio_b200_drain32.so
  SENTIO_B200_LIB=$lib timeout 600 python bench.py --workload rerank --cpu-sample 0 --no-extras > gpurun_out/ab_ce_$v.json 2> gpurun_out/ab_ce_$v.err
  echo "bench $v rc=$?" >> gpurun_out/status.txt
done
cat gpurun_out/status.txt; tail -3 gpurun_out/pytest_ce.log | cut -c1-300
for v in drain32 base drain32b base2; do tail -1 gpurun_out/ab_ce_$v.json | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); r=d['roofline']
    print('$v', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ce frac', r['cross_encoder']['frac'], d['clocks']['sm_mhz'])
except Exception as e: print('$v', 'no json', e)"; done
