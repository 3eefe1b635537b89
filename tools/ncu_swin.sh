#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
CMD="python bench.py --model swin_tiny --steps 2 --warmup 3 --no-cpu-baseline --configs none --sustain 0 --golden-state"
$CMD > gpurun_out/plain_swin.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_swin.csv $CMD > gpurun_out/ncu_swin1.log 2>&1
echo "launch list exit $?"
ncu --set full --clock-control none --import-source on -k regex:window_attention_tc -s 6 -c 3 -o gpurun_out/prof_swin_wa -f $CMD > gpurun_out/ncu_swin2.log 2>&1
echo "full exit $?"
ls -la gpurun_out | tail -5
