#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -x -q -k "gemm" > gpurun_out/gemm_tests.log 2>&1
echo "exit $?"; tail -n 3 gpurun_out/gemm_tests.log
timeout 200 python tools/gemm_bench.py deit_small 256 2 2>&1 | tee gpurun_out/gemm_bench_deit_small.log
timeout 200 python tools/gemm_bench.py vit_base 256 2 2>&1 | tee gpurun_out/gemm_bench_vit_base.log
timeout 200 python tools/gemm_bench.py vit_large 128 2 2>&1 | tee gpurun_out/gemm_bench_vit_large.log
timeout 200 python tools/gemm_bench.py deit_tiny 256 2 2>&1 | tee gpurun_out/gemm_bench_deit_tiny.log
