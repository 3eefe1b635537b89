#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
timeout 100 python tools/att_bench.py 6 256 > gpurun_out/plain_att.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 1 -c 1 -o gpurun_out/prof_att -f python tools/att_bench.py 6 256 > gpurun_out/ncu_att.log 2>&1
echo "exit $?"; tail -n 3 gpurun_out/ncu_att.log
