#!/bin/bash
# ncu capture of the stand-alone block GEMMs (tools/gemm_bench.py), variant given as $2 (default 2 = CTA-pair kernel)
export PYTHONPATH=$PWD
export GEMM_BENCH_ONCE=1
mkdir -p gpurun_out
M=${1:-deit_small}; V=${2:-2}; TAG=${3:-pair}
timeout 200 python tools/gemm_bench.py $M 256 $V > gpurun_out/plain_gemm_$TAG.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_ -c 8 -o gpurun_out/prof_gemm_$TAG -f \
   python tools/gemm_bench.py $M 256 $V > gpurun_out/ncu_gemm_$TAG.log 2>&1
echo "exit $?"; tail -n 5 gpurun_out/ncu_gemm_$TAG.log
