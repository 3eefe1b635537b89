#!/bin/bash
# usage: tools/gpu_prof.sh <tag>: ncu launch list of a 2-step bench run + one full capture of a steady-state block's kernels
export PYTHONPATH=$PWD
bash tools/gpu_launches.sh $1
bash tools/gpu_ncu.sh $1 'gemm_pair|attention_tc|layernorm' 40 9
