#!/bin/bash
# round-2 evidence for the headline workload: launch list of one bench run + one full capture of the step's kernels (one forward)
export PYTHONPATH=$PWD
export P2VIT_ATT_AUTOTUNE=0     # the engine would otherwise time both attention probability paths per layer inside the listed run (extra launches); DeiT-S keeps mode 0 anyway
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --configs none --sustain 0 --golden-state"
$CMD > gpurun_out/plain_r2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_r2.csv $CMD > gpurun_out/ncu_r2a.log 2>&1
echo "launch list exit $?"
# one forward = 89 launches; skip the eager warm-up + capture (3 forwards) and take one block's kernels + head
ncu --set full --clock-control none --import-source on -k regex:"gemm_pair|attention_tc|layernorm_pot" -s 40 -c 7 -o gpurun_out/prof_r2 -f $CMD > gpurun_out/ncu_r2b.log 2>&1
echo "full exit $?"
ls -la gpurun_out | tail -4
