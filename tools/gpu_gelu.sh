#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
echo "=== gelu tests"
timeout 300 python -m pytest tests/test_gpu_ops.py -x -q -k "gelu" > gpurun_out/gelu_tests.log 2>&1
rc=$?; echo "exit $rc"; tail -n 12 gpurun_out/gelu_tests.log
timeout 200 python tools/gemm_bench.py deit_small 256 2 2>&1 | tee gpurun_out/gemm_bench_deit_small.log
