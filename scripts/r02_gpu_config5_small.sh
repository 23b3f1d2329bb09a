#!/bin/bash
# dry run of the config-5 code path (range-seeded generators, sharded device index build, token all-gather) on 2 GPUs
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ARGS="--n-docs 2200000 --shard corpus --workload rerank --gpus 2 --steps 3 --warmup 2 --no-extras --cpu-sample 0"
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py $ARGS > gpurun_out/config5_small_g2.json 2> gpurun_out/config5_small_g2.err
echo "config5 small rc=$?"; tail -5 gpurun_out/config5_small_g2.err | cut -c1-400; tail -1 gpurun_out/config5_small_g2.json | cut -c1-600
