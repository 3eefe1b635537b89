#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
echo "=== pair tests"
timeout 300 python -m pytest tests/test_gpu_ops.py -x -q -k "pair" > gpurun_out/pair_tests.log 2>&1
rc=$?; echo "exit $rc"; tail -n 5 gpurun_out/pair_tests.log
[ $rc -ne 0 ] && exit 0
for m in "$@"; do
timeout 200 python tools/gemm_bench.py $m 256 2 > gpurun_out/gemm_bench_$m.log 2>&1
echo "exit $?"; tail -n 20 gpurun_out/gemm_bench_$m.log
P2V_PAIR_BRES=0 timeout 200 python tools/gemm_bench.py $m 256 2 2>&1 | sed 's/^/streaming-W /'
done
