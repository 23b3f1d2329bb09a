#!/bin/bash
# run 18 (1 GPU): 5-group dense test, CE tests, host-phase trace of the hybrid e2e call at batch 256, rerank bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/*.log gpurun_out/*.json gpurun_out/*.csv gpurun_out/*.txt gpurun_out/*.err
timeout 900 python -m pytest tests/test_dense_gpu.py tests/test_rerank_gpu.py tests/test_embedder_gpu.py -m gpu -q --timeout=600 -k "not full_size" > gpurun_out/pytest_sub.log 2>&1
echo "pytest_sub rc=$?" > gpurun_out/status.txt
SENTIO_B200_TRACE=1 timeout 600 python bench.py --workload hybrid --steps 6 --warmup 3 --batch 256 --cpu-sample 0 > gpurun_out/bench_hybrid_b256.json 2> gpurun_out/bench_hybrid_b256.err
echo "bench hybrid b256 rc=$?" >> gpurun_out/status.txt
SENTIO_B200_TRACE=1 timeout 600 python bench.py --workload hybrid --steps 6 --warmup 3 --batch 128 --cpu-sample 0 > gpurun_out/bench_hybrid_b128.json 2> gpurun_out/bench_hybrid_b128.err
echo "bench hybrid b128 rc=$?" >> gpurun_out/status.txt
timeout 900 python bench.py --workload rerank --steps 5 --warmup 3 --cpu-sample 0 > gpurun_out/bench_rerank.json 2> gpurun_out/bench_rerank.err
echo "bench rerank rc=$?" >> gpurun_out/status.txt
tail -4 gpurun_out/pytest_sub.log | cut -c1-200; cat gpurun_out/status.txt; grep "trace" gpurun_out/bench_hybrid_b256.json | tail -7; grep trace gpurun_out/bench_hybrid_b128.json | tail -4; for f in hybrid_b256 hybrid_b128 rerank; do tail -1 gpurun_out/bench_$f.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$f', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],3))"; done
