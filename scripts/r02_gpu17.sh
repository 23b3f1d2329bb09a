#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/*.log gpurun_out/*.json gpurun_out/*.csv gpurun_out/status*.txt gpurun_out/*.err
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?" > gpurun_out/status.txt
timeout 2400 python -m pytest tests -m gpu -x -q --timeout=1200 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest_gpu rc=$?" >> gpurun_out/status.txt
timeout 1500 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
echo "bench full rc=$?" >> gpurun_out/status.txt
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
echo "bench reference rc=$?" >> gpurun_out/status.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_dense.csv python bench.py --steps 1 --warmup 1 --inner 2 --cpu-sample 0 --no-extras > gpurun_out/ncu_launch_dense.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_hybrid.csv python bench.py --workload hybrid --steps 1 --warmup 1 --inner 1 --cpu-sample 0 --no-extras > gpurun_out/ncu_launch_hybrid.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv --log-file gpurun_out/launches_rerank.csv python bench.py --workload rerank --steps 1 --warmup 1 --inner 1 --cpu-sample 0 --no-extras > gpurun_out/ncu_launch_rerank.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k "regex:ce_gemm_ws_kernel|ce_attention_mma|ce_ln_kernel" -s 12 -c 7 -o gpurun_out/prof_ce_final python bench.py --workload rerank --steps 1 --warmup 1 --inner 1 --batch 16 --cpu-sample 0 --no-extras > gpurun_out/ncu_full_ce.log 2>&1
echo "ncu full ce rc=$?" >> gpurun_out/status.txt
tail -2 gpurun_out/smoke.log | cut -c1-200; tail -4 gpurun_out/pytest_gpu.log | cut -c1-300; cat gpurun_out/status.txt
for f in full reference; do tail -1 gpurun_out/bench_$f.json | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read())
    r=d.get('roofline') or {}
    print('$f', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms/step', round(d['ms_per_step'],3), 'frac', r.get('frac'), 'avg_ms', r.get('avg_launch_ms'), r.get('other_dense_stages_ms_per_batch'), d.get('clocks'))
    for k,v in (d.get('workloads') or {}).items(): print('  ', k, {a:(round(b,1) if isinstance(b,float) else b) for a,b in v.items() if a in ('value','ms_per_step','cpu_reference_qps','gpu_batch_qps','gpu_retrieve_one_by_one_qps','ids_identical_to_reference_path','error','timed_region_s')}, 'e2e', (v.get('e2e') or {}).get('value'))
    print('  latency', d.get('latency_b1'))
except Exception as e: print('$f', 'no json', e)"; done
