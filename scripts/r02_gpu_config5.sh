#!/bin/bash
# BASELINE config 5: 8 M docs partitioned over N GPUs (contiguous doc ranges), hybrid + cross-encoder rerank, one NCCL
# all-gather of per-shard top-k per batch.  usage: r02_gpu_config5.sh N
cd "$(dirname "$0")/.."
N=${1:-8}
mkdir -p gpurun_out
ARGS="--n-docs 8000000 --shard corpus --workload rerank --gpus $N --steps 5 --warmup 2 --no-extras --cpu-sample 0"
if [ "$N" = "1" ]; then
  timeout 2400 python bench.py $ARGS > gpurun_out/config5_g$N.json 2> gpurun_out/config5_g$N.err
else
  timeout 2400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py $ARGS > gpurun_out/config5_g$N.json 2> gpurun_out/config5_g$N.err
fi
echo "config5 N=$N rc=$?" > gpurun_out/config5_status_$N.txt
cat gpurun_out/config5_status_$N.txt; tail -3 gpurun_out/config5_g$N.err | cut -c1-300
tail -1 gpurun_out/config5_g$N.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('config5 N=$N', round(d['value'],1), 'q/s  e2e', round(d['e2e']['value'],1), 'ms/step', round(d['ms_per_step'],2), d['config']['multi_gpu'], 'gen_s', d['corpus_gen_s'])
print('  scan frac', r.get('frac'), 'bm25', (r.get('bm25') or {}).get('share_of_step'), 'ce', (r.get('cross_encoder') or {}).get('frac'), (r.get('cross_encoder') or {}).get('share_of_step'))"
