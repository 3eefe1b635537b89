#!/bin/bash
# usage: tools/gpu_r1l.sh <tag>   tests + smoke + bench, then launch list and one full capture of each hot kernel
export PYTHONPATH=$PWD
mkdir -p gpurun_out
TAG=$1
echo "=== tests"
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/tests.log 2>&1
echo "exit $?"; tail -n 4 gpurun_out/tests.log
echo "=== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "exit $?"; tail -n 2 gpurun_out/smoke.log
echo "=== bench"
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.log 2>&1
echo "exit $?"; tail -n 1 gpurun_out/bench_$TAG.log | cut -c1-3000
CMD="python bench.py --steps 2 --warmup 3 --golden-state --no-cpu-baseline"
echo "=== launches"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'gemm_|attention|layernorm|patchify|fill_cls' -c 360 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "exit $?"
echo "=== full"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'gemm_pair|attention_tc|layernorm' -s 40 -c 9 -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_full.log 2>&1
echo "exit $?"; tail -2 gpurun_out/ncu_full.log | cut -c1-200; ls -la gpurun_out/*.ncu-rep
