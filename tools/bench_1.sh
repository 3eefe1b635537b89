#!/bin/bash
# the driver's 1-GPU lines: smoke, the default bench line and the reference arm -> gpurun_out/
export PYTHONPATH=$PWD
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench exit $?"; cut -c1-300 gpurun_out/bench_default.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference exit $?"; cut -c1-400 gpurun_out/bench_reference.json
