#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout ${TMO:-900} "$@" > gpurun_out/$name.log 2>&1; echo "exit $?"; tail -n ${TAILN:-6} gpurun_out/$name.log | cut -c1-1800; }
run bench_vitb python bench.py --model vit_base --batch 256 --steps 10 --warmup 3 --no-cpu-baseline
run bench_vitl python bench.py --model vit_large --batch 128 --steps 5 --warmup 3 --no-cpu-baseline --bits mixed
run bench_tiny python bench.py --model deit_tiny --batch 256 --steps 20 --warmup 3 --no-cpu-baseline
