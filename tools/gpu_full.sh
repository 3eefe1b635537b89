#!/bin/bash
# full GPU check: all gpu tests, smoke, bench (DeiT-S) -> gpurun_out/
export PYTHONPATH=$PWD
mkdir -p gpurun_out
echo "=== tests"
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/tests.log 2>&1
echo "exit $?"; tail -n 6 gpurun_out/tests.log
echo "=== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "exit $?"; tail -n 3 gpurun_out/smoke.log
echo "=== bench"
timeout 600 python bench.py --steps 20 --warmup 3 ${BENCH_ARGS} > gpurun_out/bench.log 2>&1
echo "exit $?"; tail -n 1 gpurun_out/bench.log | cut -c1-2500
