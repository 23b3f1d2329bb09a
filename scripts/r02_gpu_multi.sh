#!/bin/bash
# r02 N-GPU session: sharded == single checks, then the default bench line (headline replicas + partitioned leg)
cd "$(dirname "$0")/.."
N=${1:-2}
mkdir -p gpurun_out
rm -f gpurun_out/*_g$N.json gpurun_out/*_g$N.err gpurun_out/multi_*_$N.*
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29533 scripts/check_multigpu.py > gpurun_out/multi_check_$N.log 2>&1
echo "check rc=$?" > gpurun_out/multi_status_$N.txt
timeout 900 $TR --master-port 29534 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_default_g$N.json 2> gpurun_out/bench_default_g$N.err
echo "bench default rc=$?" >> gpurun_out/multi_status_$N.txt
grep -n "==\|ok\|Error" gpurun_out/multi_check_$N.log | tail -12; cat gpurun_out/multi_status_$N.txt; tail -3 gpurun_out/bench_default_g$N.err
tail -1 gpurun_out/bench_default_g$N.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('headline', round(d['value']), 'e2e', round(d['e2e']['value']), d['config']['multi_gpu'])
for k,v in (d.get('workloads') or {}).items(): print(' ', k, round(v.get('value',0)), 'e2e', round((v.get('e2e') or {}).get('value',0)))
p=d.get('partitioned')
if p:
    print(' partitioned dense', round(p['dense']['value']), 'e2e', round(p['dense']['e2e']['value']), {k:round(v,1) for k,v in p['per_batch_stage_us_rank0'].items()})
    if 'hybrid' in p: print(' partitioned hybrid', round(p['hybrid']['value']), 'e2e', round(p['hybrid']['e2e']['value']))
"
