#!/bin/bash
# r02 run 11b (1 GPU): cross-encoder variants -- rotated attention row tiles + trimmed softmax (default build) vs the
# unrotated mapping (libsentio_b200_norot.so) vs the tanh-fit GELU epilogue (libsentio_b200_gelut.so); long-window tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/*.log gpurun_out/*.json gpurun_out/*.csv gpurun_out/status*.txt gpurun_out/*.err
timeout 900 python -m pytest tests/test_rerank_gpu.py tests/test_embedder_gpu.py -m gpu -x -q --timeout=600 > gpurun_out/pytest_ce.log 2>&1
echo "pytest_ce rc=$?" > gpurun_out/status.txt
SENTIO_B200_LIB=$PWD/sentio_b200/libsentio_b200_gelut.so timeout 900 python -m pytest tests/test_rerank_gpu.py -m gpu -x -q --timeout=600 > gpurun_out/pytest_ce_gelut.log 2>&1
echo "pytest_ce_gelut rc=$?" >> gpurun_out/status.txt
for v in base norot gelut base2; do
  lib=$PWD/sentio_b200/libsentio_b200.so
  [ $v = norot ] && lib=$PWD/sentio_b200/libsentio_b200_norot.so
  [ $v = gelut ] && lib=$PWD/sentio_b200/libsentio_b200_gelut.so
  SENTIO_B200_LIB=$lib timeout 600 python bench.py --workload rerank --cpu-sample 0 --no-extras > gpurun_out/ab_ce_$v.json 2> gpurun_out/ab_ce_$v.err
  echo "bench $v rc=$?" >> gpurun_out/status.txt
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv --log-file gpurun_out/launches_rerank.csv python bench.py --workload rerank --steps 1 --warmup 1 --inner 1 --cpu-sample 0 --no-extras > gpurun_out/ncu_launch_rerank.log 2>&1
SENTIO_B200_LIB=$PWD/sentio_b200/libsentio_b200_gelut.so timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv --log-file gpurun_out/launches_rerank_gelut.csv python bench.py --workload rerank --steps 1 --warmup 1 --inner 1 --cpu-sample 0 --no-extras > gpurun_out/ncu_launch_rerank_gelut.log 2>&1
cat gpurun_out/status.txt; tail -3 gpurun_out/pytest_ce.log | cut -c1-300; tail -3 gpurun_out/pytest_ce_gelut.log | cut -c1-300
for v in base norot gelut base2; do tail -1 gpurun_out/ab_ce_$v.json | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); r=d['roofline']
    print('$v', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ce frac', (r.get('cross_encoder') or r).get('frac'), d['clocks']['sm_mhz'])
except Exception as e: print('$v', 'no json', e)"; done
