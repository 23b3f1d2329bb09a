#!/bin/bash
# N-GPU session: default (auto) layout and an explicit 2-D layout (C corpus shards x N/C query groups)
cd "$(dirname "$0")/.."
N=${1:-4}
C=${2:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29534 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_dense_auto_g$N.json 2> gpurun_out/bench_dense_auto_g$N.err
echo "bench dense auto rc=$?" > gpurun_out/multi_status_$N.txt
timeout 400 $TR --master-port 29537 bench.py --gpus $N --steps 20 --warmup 3 --corpus-shards $C > gpurun_out/bench_dense_c${C}_g$N.json 2> gpurun_out/bench_dense_c${C}_g$N.err
echo "bench dense C=$C rc=$?" >> gpurun_out/multi_status_$N.txt
cat gpurun_out/multi_status_$N.txt
for f in dense_auto dense_c${C}; do tail -1 gpurun_out/bench_${f}_g$N.json | cut -c1-900; echo; tail -2 gpurun_out/bench_${f}_g$N.err; done
