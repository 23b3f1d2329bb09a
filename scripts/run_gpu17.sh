#!/bin/bash
# run 17 (1 GPU): attention with multi-head CTAs + cp.async double buffering; batch-256 benches
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/*.log gpurun_out/*.json gpurun_out/*.csv gpurun_out/*.txt gpurun_out/*.err
timeout 900 python -m pytest tests/test_rerank_gpu.py tests/test_embedder_gpu.py tests/test_hybrid_e2e.py -m gpu -q --timeout=600 > gpurun_out/pytest_ce.log 2>&1
echo "pytest_ce rc=$?" > gpurun_out/status.txt
timeout 900 python bench.py --workload rerank --steps 5 --warmup 3 --cpu-sample 0 > gpurun_out/bench_rerank.json 2> gpurun_out/bench_rerank.err
echo "bench rerank rc=$?" >> gpurun_out/status.txt
timeout 400 python bench.py --steps 10 --warmup 3 --batch 256 --cpu-sample 0 > gpurun_out/bench_dense_b256.json 2> gpurun_out/bench_dense_b256.err
echo "bench dense b256 rc=$?" >> gpurun_out/status.txt
timeout 600 python bench.py --workload hybrid --steps 6 --warmup 3 --batch 256 --cpu-sample 0 > gpurun_out/bench_hybrid_b256.json 2> gpurun_out/bench_hybrid_b256.err
echo "bench hybrid b256 rc=$?" >> gpurun_out/status.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 170 --csv --log-file gpurun_out/launches_rerank.csv python bench.py --workload rerank --steps 1 --warmup 1 --batch 16 --cpu-sample 0 > gpurun_out/ncu_launch_rerank.log 2>&1
echo "ncu launches rerank rc=$?" >> gpurun_out/status.txt
tail -4 gpurun_out/pytest_ce.log | cut -c1-200; cat gpurun_out/status.txt; for f in rerank dense_b256 hybrid_b256; do python - <<PY
import json
d=json.load(open("gpurun_out/bench_$f.json"))
print("$f", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"],3), d["roofline"].get("cross_encoder",{}).get("frac"))
PY
done; tail -2 gpurun_out/bench_rerank.err
