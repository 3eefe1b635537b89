#!/bin/bash
# Swin-T: launch list of one bench run (the per-kernel shares of the window-attention / LayerNorm / GEMM families)
export PYTHONPATH=$PWD
mkdir -p gpurun_out
CMD="python bench.py --model swin_tiny --steps 2 --warmup 3 --no-cpu-baseline --configs none --sustain 0 --golden-state"
$CMD > gpurun_out/plain_swin.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_swin.csv $CMD > gpurun_out/ncu_swin1.log 2>&1
echo "launch list exit $?"
