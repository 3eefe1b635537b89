#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
echo "=== attention tests"
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -k "attention" > gpurun_out/att_tests.log 2>&1
rc=$?; echo "exit $rc"; tail -n 6 gpurun_out/att_tests.log
[ $rc -ne 0 ] && exit 0
timeout 100 python tools/att_bench.py 6 256 2>&1 | tee gpurun_out/att_bench.log
timeout 100 python tools/att_bench.py 12 256 2>&1 | tee -a gpurun_out/att_bench.log
