#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?"; tail -n 40 gpurun_out/$name.log; }
run model python -m pytest tests/test_gpu_model.py -q -m gpu -p no:cacheprovider
run diag python tools/diag_parity.py deit_tiny 8 8
