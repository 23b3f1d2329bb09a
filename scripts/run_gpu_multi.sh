#!/bin/bash
# multi-GPU session (gpurun --gpus N): sharded == single check, then bench at N ranks
cd "$(dirname "$0")/.."
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 scripts/check_multigpu.py > gpurun_out/multigpu_check_$N.log 2>&1
echo "check rc=$?" > gpurun_out/multi_status_$N.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_dense_g$N.json 2> gpurun_out/bench_dense_g$N.err
echo "bench dense rc=$?" >> gpurun_out/multi_status_$N.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus $N --workload hybrid --steps 10 --warmup 3 > gpurun_out/bench_hybrid_g$N.json 2> gpurun_out/bench_hybrid_g$N.err
echo "bench hybrid rc=$?" >> gpurun_out/multi_status_$N.txt
tail -5 gpurun_out/multigpu_check_$N.log; cat gpurun_out/multi_status_$N.txt; tail -1 gpurun_out/bench_dense_g$N.json; tail -1 gpurun_out/bench_hybrid_g$N.json; tail -3 gpurun_out/bench_dense_g$N.err
if [ "${2:-}" = "big" ]; then
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29536 bench.py --gpus $N --n-docs 8000000 --steps 20 --warmup 3 > gpurun_out/bench_dense8m_g$N.json 2> gpurun_out/bench_dense8m_g$N.err
  echo "bench dense 8M rc=$?" >> gpurun_out/multi_status_$N.txt
  tail -1 gpurun_out/bench_dense8m_g$N.json; tail -3 gpurun_out/bench_dense8m_g$N.err
fi
