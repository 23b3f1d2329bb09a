#!/bin/bash
# the driver's scaling command at N GPUs: default bench line (replicated headline + workloads + partitioned leg)
cd "$(dirname "$0")/.."
N=${1:-4}
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_default_g$N.json 2> gpurun_out/bench_default_g$N.err
echo "bench default N=$N rc=$?"; tail -2 gpurun_out/bench_default_g$N.err | cut -c1-300
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
