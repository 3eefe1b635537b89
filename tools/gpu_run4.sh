#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout 1200 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?"; tail -n ${TAILN:-12} gpurun_out/$name.log; }
run ops python -m pytest tests/test_gpu_ops.py -q -m gpu -p no:cacheprovider
run model python -m pytest tests/test_gpu_model.py -q -m gpu -p no:cacheprovider
TAILN=3 run bench python bench.py --steps 20 --warmup 3 --golden-state --no-cpu-baseline
