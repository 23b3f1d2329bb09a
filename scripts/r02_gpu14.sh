#!/bin/bash
# r02 run 15 (1 GPU): cross-encoder GEMM variants -- 4 accumulator stages in tensor memory (acc4), + register bias for the
# weight-stationary bias / GELU epilogues (acc4b) -- against the default build; GEMM + model tests under acc4b
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/*.log gpurun_out/*.json gpurun_out/*.csv gpurun_out/status*.txt gpurun_out/*.err
for v in base acc4 acc4b base2; do
  lib=$PWD/sentio_b200/libsentio_b200.so
  [ $v = acc4 ] && lib=$PWD/sentio_b200/libsentio_b200_acc4.so
  [ $v = acc4b ] && lib=$PWD/sentio_b200/libsentio_b200_acc4b.so
  SENTIO_B200_LIB=$lib timeout 600 python bench.py --workload rerank --cpu-sample 0 --no-extras > gpurun_out/ab_ce_$v.json 2> gpurun_out/ab_ce_$v.err
  echo "bench $v rc=$?" >> gpurun_out/status.txt
done
SENTIO_B200_LIB=$PWD/sentio_b200/libsentio_b200_acc4b.so timeout 600 python -m pytest tests/test_rerank_gpu.py tests/test_embedder_gpu.py -m gpu -x -q --timeout=600 > gpurun_out/pytest_ce_acc4b.log 2>&1
echo "pytest_ce_acc4b rc=$?" >> gpurun_out/status.txt
SENTIO_B200_LIB=$PWD/sentio_b200/libsentio_b200_acc4b.so timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv --log-file gpurun_out/launches_rerank_acc4b.csv python bench.py --workload rerank --steps 1 --warmup 1 --inner 1 --cpu-sample 0 --no-extras > gpurun_out/ncu_launch_rerank_acc4b.log 2>&1
SENTIO_B200_LIB=$PWD/sentio_b200/libsentio_b200_acc4.so timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv --log-file gpurun_out/launches_rerank_acc4.csv python bench.py --workload rerank --steps 1 --warmup 1 --inner 1 --cpu-sample 0 --no-extras > gpurun_out/ncu_launch_rerank_acc4.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv --log-file gpurun_out/launches_rerank_base.csv python bench.py --workload rerank --steps 1 --warmup 1 --inner 1 --cpu-sample 0 --no-extras > gpurun_out/ncu_launch_rerank_base.log 2>&1
cat gpurun_out/status.txt; tail -3 gpurun_out/pytest_ce_acc4b.log | cut -c1-300
for v in base acc4 acc4b base2; do tail -1 gpurun_out/ab_ce_$v.json | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); r=d['roofline']
    print('$v', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ce frac', r['cross_encoder']['frac'], d['clocks']['sm_mhz'])
except Exception as e: print('$v', 'no json', e)"; done
