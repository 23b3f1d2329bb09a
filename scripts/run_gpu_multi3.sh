#!/bin/bash
# reduced multi-GPU session (gpurun --gpus N): dense bench in both shard modes at N ranks
cd "$(dirname "$0")/.."
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29534 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_dense_g$N.json 2> gpurun_out/bench_dense_g$N.err
echo "bench dense rc=$?" > gpurun_out/multi_status_$N.txt
timeout 600 $TR --master-port 29537 bench.py --gpus $N --steps 20 --warmup 3 --shard queries > gpurun_out/bench_dense_replicas_g$N.json 2> gpurun_out/bench_dense_replicas_g$N.err
echo "bench dense replicas rc=$?" >> gpurun_out/multi_status_$N.txt
cat gpurun_out/multi_status_$N.txt
for f in dense dense_replicas; do tail -1 gpurun_out/bench_${f}_g$N.json | cut -c1-420; echo; tail -2 gpurun_out/bench_${f}_g$N.err; done
