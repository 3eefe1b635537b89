#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
python -m pytest tests/test_gpu_ops.py -q -k "layernorm" > gpurun_out/ab_ln_tests.log 2>&1; echo "tests rc $?"; tail -2 gpurun_out/ab_ln_tests.log
for c in 384 192 768; do python tools/ln_bench.py $c; done 2>&1 | tee gpurun_out/ab_ln.log
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --configs none --sustain 0 --golden-state > gpurun_out/ab_bench.json 2>gpurun_out/ab_bench.err; echo "bench rc $?"
python - <<PY
import json
d=json.load(open('gpurun_out/ab_bench.json'))
print(d['value'], d['ms_per_step'], d['roofline']['device_ms_per_step_by_family'])
PY
