#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
for m in "$@"; do
timeout 200 python tools/gemm_bench.py $m 256 > gpurun_out/gemm_bench_$m.log 2>&1
echo "exit $?"; tail -n 20 gpurun_out/gemm_bench_$m.log
done
