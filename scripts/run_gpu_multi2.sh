#!/bin/bash
# multi-GPU session (gpurun --gpus N): sharded == single check, then the weak-scaling benches at N ranks
cd "$(dirname "$0")/.."
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29533 scripts/check_multigpu.py > gpurun_out/multigpu_check_$N.log 2>&1
echo "check rc=$?" > gpurun_out/multi_status_$N.txt
timeout 600 $TR --master-port 29534 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_dense_g$N.json 2> gpurun_out/bench_dense_g$N.err
echo "bench dense rc=$?" >> gpurun_out/multi_status_$N.txt
timeout 600 $TR --master-port 29537 bench.py --gpus $N --steps 20 --warmup 3 --shard queries > gpurun_out/bench_dense_replicas_g$N.json 2> gpurun_out/bench_dense_replicas_g$N.err
echo "bench dense replicas rc=$?" >> gpurun_out/multi_status_$N.txt
timeout 900 $TR --master-port 29535 bench.py --gpus $N --workload hybrid --steps 10 --warmup 3 > gpurun_out/bench_hybrid_g$N.json 2> gpurun_out/bench_hybrid_g$N.err
echo "bench hybrid rc=$?" >> gpurun_out/multi_status_$N.txt
timeout 900 $TR --master-port 29536 bench.py --gpus $N --workload rerank --steps 5 --warmup 3 > gpurun_out/bench_rerank_g$N.json 2> gpurun_out/bench_rerank_g$N.err
echo "bench rerank rc=$?" >> gpurun_out/multi_status_$N.txt
tail -6 gpurun_out/multigpu_check_$N.log; cat gpurun_out/multi_status_$N.txt
for f in dense dense_replicas hybrid rerank; do tail -1 gpurun_out/bench_${f}_g$N.json | cut -c1-420; echo; tail -2 gpurun_out/bench_${f}_g$N.err; done
