#!/bin/bash
# usage: tools/gpu_ncu.sh <tag> <kernel regex> <skip> <count>
mkdir -p gpurun_out
TAG=$1; RE=$2; SKIP=${3:-30}; CNT=${4:-14}
CMD="python bench.py --steps 2 --warmup 3 --golden-state --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/plain_for_ncu.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"$RE" -s $SKIP -c $CNT -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_full.log 2>&1
echo "exit $?"; tail -3 gpurun_out/ncu_full.log | cut -c1-300; ls -la gpurun_out/*.ncu-rep
