#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout ${TMO:-900} "$@" > gpurun_out/$name.log 2>&1; echo "exit $?"; tail -n ${TAILN:-40} gpurun_out/$name.log; }
run model python -m pytest ${TESTFILE:-tests/test_gpu_model.py} -q -m gpu -p no:cacheprovider "$@"
