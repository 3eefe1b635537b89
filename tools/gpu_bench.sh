#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout ${TMO:-600} "$@" > gpurun_out/$name.log 2>&1; echo "exit $?"; tail -n ${TAILN:-15} gpurun_out/$name.log; }
run gemm python -m pytest tests/test_gpu_ops.py -q -m gpu -p no:cacheprovider -k "gemm" -x
TAILN=3 run bench python bench.py --steps 20 --warmup 3 --golden-state --no-cpu-baseline
