#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
echo "=== bench x$N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n$N.log 2>&1
echo "exit $?"; tail -n 3 gpurun_out/bench_n$N.log | cut -c1-1500
echo "=== reference arm x$N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_ref_n$N.log 2>&1
echo "exit $?"; tail -n 2 gpurun_out/bench_ref_n$N.log | cut -c1-600
