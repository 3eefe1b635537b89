#!/bin/bash
export PYTHONPATH=$PWD
export P2V_LIB=$PWD/p2vit_b200/csrc/libp2vit_b200_trace.so
mkdir -p gpurun_out
for k in "$@"; do
kind=${k%%:*}; d=${k##*:}; [ "$d" = "$k" ] && d=384
timeout 100 python tools/pair_trace.py $kind $d > gpurun_out/trace_${kind}_$d.log 2>&1; echo "exit $?"
done
