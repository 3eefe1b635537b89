#!/bin/bash
# one gpurun call = this script with a list of named jobs: tools/gpu_job.sh tests peak bench:deit_small ...
export PYTHONPATH=$PWD
mkdir -p gpurun_out
for job in "$@"; do
  name=${job%%:*}; arg=${job#*:}
  echo "=== $job"
  case $name in
    tests) timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider ${arg:+-k "$arg"} > gpurun_out/tests.log 2>&1; echo "exit $?"; tail -n 25 gpurun_out/tests.log;;
    testsall) timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/tests.log 2>&1; echo "exit $?"; tail -n 40 gpurun_out/tests.log;;
    smoke) timeout 300 python __graft_entry__.py smoke 2>&1 | tail -n 3;;
    peak) timeout 300 python tools/int8_peak.py gpurun_out/int8_peak.json 2>&1 | tail -n 2;;
    bench) timeout 900 python bench.py --model $arg --steps 20 --warmup 3 ${BENCH_FLAGS} > gpurun_out/bench_$arg.json 2> gpurun_out/bench_$arg.err; echo "exit $?"; cut -c1-600 gpurun_out/bench_$arg.json; tail -n 3 gpurun_out/bench_$arg.err;;
    benchdef) timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "exit $?"; cut -c1-1500 gpurun_out/bench_default.json; tail -n 3 gpurun_out/bench_default.err;;
    py) timeout 900 python $arg 2>&1 | tail -n 40;;
    sh) timeout 1200 bash $arg 2>&1 | tail -n 60;;
  esac
done
