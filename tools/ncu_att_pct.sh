#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
CMD="python bench.py --model vit_base --method percentile --steps 2 --warmup 3 --no-cpu-baseline --configs none --sustain 0"
ncu --set full --clock-control none --import-source on -k regex:"attention_tc" -s 30 -c 1 -o gpurun_out/prof_att_pct -f $CMD > gpurun_out/ncu_att_pct.log 2>&1
echo "exit $?"
