#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
for c in 384 768 192; do
  python tools/ln_bench.py $c
  P2V_LN_LPR=32 python tools/ln_bench.py $c | sed 's/^/LPR32 /'
done 2>&1 | tee gpurun_out/ln_bench.log
