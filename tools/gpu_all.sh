#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
echo "=== ops tests"
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q > gpurun_out/ops_tests.log 2>&1
rc=$?; echo "exit $rc"; tail -n 4 gpurun_out/ops_tests.log
[ $rc -ne 0 ] && exit 0
timeout 100 python tools/att_bench.py 6 256 2>&1 | tee gpurun_out/att_bench.log
timeout 200 python tools/gemm_bench.py deit_small 256 2 2>&1 | tee gpurun_out/gemm_bench_deit_small.log
timeout 100 python tools/ln_bench.py 384 2>&1 | tee gpurun_out/ln_bench.log
