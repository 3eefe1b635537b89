#!/bin/bash
# A/B of programmatic dependent launch (common.cuh): GPU tests with it on, GPU tests in the default mode (captured launches only) and with P2VIT_PDL=1 (every launch), then the bench with P2VIT_PDL=0 and the default.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/tests.log
P2VIT_PDL=1 python -m pytest tests -m gpu -x -q > gpurun_out/tests_pdl1.log 2>&1; echo "tests (PDL) rc=$?"; tail -2 gpurun_out/tests_pdl1.log
P2VIT_PDL=0 python bench.py --no-cpu-baseline "$@" > gpurun_out/bench_pdl0.json 2> gpurun_out/bench_pdl0.err; echo "rc=$?"
python bench.py --no-cpu-baseline "$@" > gpurun_out/bench_pdl1.json 2> gpurun_out/bench_pdl1.err; echo "rc=$?"
for f in gpurun_out/bench_pdl0.json gpurun_out/bench_pdl1.json; do
  python -c "
import json,sys
d=json.load(open(sys.argv[1]))
print(sys.argv[1], round(d['value']), round(d['e2e']['value']), round(d['e2e_u8']['value']) if d.get('e2e_u8') else None, d['ms_per_step'])" $f
done
