#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
python -m pytest tests/test_gpu_ops.py -q -k "layernorm" > gpurun_out/ab_ln_tests.log 2>&1; echo "ln tests rc $?"; tail -2 gpurun_out/ab_ln_tests.log
python -m pytest tests/test_gpu_model.py tests/test_gpu_swin.py -q -x > gpurun_out/ab_ln_model.log 2>&1; echo "model tests rc $?"; tail -2 gpurun_out/ab_ln_model.log
(python tools/ln_bench.py 96 802816; python tools/ln_bench.py 192 200704; python tools/ln_bench.py 96 50432; python tools/ln_bench.py 192 50432; python tools/ln_bench.py 384 50432) 2>&1 | tee gpurun_out/ab_ln.log
for spec in "swin_tiny minmax" "deit_tiny minmax"; do
  set -- $spec
  python bench.py --model $1 --method $2 --steps 10 --warmup 3 --no-cpu-baseline --configs none --sustain 0 > gpurun_out/ln2_$1_$2.json 2>gpurun_out/ln2.err || tail -3 gpurun_out/ln2.err
  python - <<PY
import json
d=json.load(open('gpurun_out/ln2_$1_$2.json'))
r=d['roofline']
print('$1 $2', round(d['value']), d['ms_per_step'], r.get('device_ms_per_step_by_family'))
PY
done
