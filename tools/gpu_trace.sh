#!/bin/bash
export PYTHONPATH=$PWD
export P2V_LIB=$PWD/p2vit_b200/csrc/libp2vit_b200_trace.so
mkdir -p gpurun_out
for k in "$@"; do
timeout 100 python tools/pair_trace.py $k > gpurun_out/trace_$k.log 2>&1; echo "exit $?"
done
