#!/bin/bash
# r02 run 14 (1 GPU): BM25 range kernel with L1-prefetched posting chunks, streaming dense-row loads and packed per-term
# state vs the previous kernel (libsentio_b200_bm25old.so); cross-encoder with the tanh-fit GELU as the default build
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/*.log gpurun_out/*.json gpurun_out/*.csv gpurun_out/status*.txt gpurun_out/*.err
timeout 900 python -m pytest tests/test_bm25_gpu.py tests/test_hybrid_e2e.py tests/test_rerank_gpu.py tests/test_embedder_gpu.py -m gpu -x -q --timeout=600 > gpurun_out/pytest_part.log 2>&1
echo "pytest_part rc=$?" > gpurun_out/status.txt
for v in old new old2 new2; do
  lib=$PWD/sentio_b200/libsentio_b200.so
  [ ${v:0:3} = old ] && lib=$PWD/sentio_b200/libsentio_b200_bm25old.so
  SENTIO_B200_LIB=$lib timeout 600 python bench.py --workload hybrid --cpu-sample 0 --no-extras > gpurun_out/ab_bm25_$v.json 2> gpurun_out/ab_bm25_$v.err
  echo "bench $v rc=$?" >> gpurun_out/status.txt
done
timeout 600 python bench.py --workload rerank --cpu-sample 0 --no-extras > gpurun_out/bench_rerank.json 2> gpurun_out/bench_rerank.err
echo "bench rerank rc=$?" >> gpurun_out/status.txt
for v in acc4 acc4b; do
  SENTIO_B200_LIB=$PWD/sentio_b200/libsentio_b200_$v.so timeout 600 python bench.py --workload rerank --cpu-sample 0 --no-extras > gpurun_out/ab_ce_$v.json 2> gpurun_out/ab_ce_$v.err
  echo "bench $v rc=$?" >> gpurun_out/status.txt
  tail -1 gpurun_out/ab_ce_$v.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$v rerank', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['roofline']['cross_encoder']['frac'])" >> gpurun_out/status.txt
done
SENTIO_B200_LIB=$PWD/sentio_b200/libsentio_b200_acc4b.so timeout 600 python -m pytest tests/test_rerank_gpu.py -m gpu -x -q --timeout=600 > gpurun_out/pytest_ce_acc4b.log 2>&1
echo "pytest_ce_acc4b rc=$?" >> gpurun_out/status.txt
SENTIO_B200_LIB=$PWD/sentio_b200/libsentio_b200_acc4b.so timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv --log-file gpurun_out/launches_rerank_acc4b.csv python bench.py --workload rerank --steps 1 --warmup 1 --inner 1 --cpu-sample 0 --no-extras > gpurun_out/ncu_launch_rerank_acc4b.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_hybrid.csv python bench.py --workload hybrid --steps 1 --warmup 1 --inner 1 --cpu-sample 0 --no-extras > gpurun_out/ncu_launch_hybrid.log 2>&1
cat gpurun_out/status.txt; tail -3 gpurun_out/pytest_part.log | cut -c1-300
for v in old new old2 new2; do tail -1 gpurun_out/ab_bm25_$v.json | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); r=d['roofline']['bm25']
    print('$v', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'bm25 ms', round(r['ms_total'],1), 'share', round(r['share_of_step'],3), d['clocks']['sm_mhz'])
except Exception as e: print('$v', 'no json', e)"; done
tail -1 gpurun_out/bench_rerank.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('rerank', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['roofline']['cross_encoder']['frac'])"
