#!/bin/bash
# tools/bench_n.sh N : the driver's multi-GPU launch of bench.py (torchrun, one rank per GPU) -> gpurun_out/bench_n$N.json
N=${1:-2}
export PYTHONPATH=$PWD
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "exit $?"; cut -c1-400 gpurun_out/bench_n$N.json; tail -n 5 gpurun_out/bench_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/bench_ref_n$N.json 2> gpurun_out/bench_ref_n$N.err
echo "exit $?"; cut -c1-300 gpurun_out/bench_ref_n$N.json
