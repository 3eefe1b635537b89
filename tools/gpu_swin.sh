#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout ${TMO:-900} "$@" > gpurun_out/$name.log 2>&1; echo "exit $?"; tail -n ${TAILN:-6} gpurun_out/$name.log | cut -c1-2500; }
run bench_swint python bench.py --model swin_tiny --batch 256 --steps 10 --warmup 3
