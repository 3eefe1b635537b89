#!/bin/bash
# usage: tools/gpu_launches.sh <tag>
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --golden-state --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/plain_for_ncu.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'gemm_|attention|layernorm|patchify|fill_cls' -c 360 --csv --log-file gpurun_out/launches_$1.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "exit $?"; tail -2 gpurun_out/ncu_launches.log | cut -c1-200
