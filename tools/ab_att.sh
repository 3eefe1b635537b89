#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
P2V_ATT_EXACT=1 python -m pytest tests/test_gpu_swin.py -q > gpurun_out/ab_swin_tests.log 2>&1; echo "swin tests (exact) rc $?"; tail -2 gpurun_out/ab_swin_tests.log
for x in 0 1; do
  P2V_ATT_EXACT=$x python bench.py --model swin_tiny --steps 10 --warmup 3 --no-cpu-baseline --configs none --sustain 0 > gpurun_out/ex_swin.json 2>gpurun_out/ex.err || tail -3 gpurun_out/ex.err
  python - <<PY
import json
d=json.load(open('gpurun_out/ex_swin.json'))
r=d['roofline']
print('swin_tiny exact=$x', round(d['value']), d['ms_per_step'], r['device_ms_per_step_by_family'])
PY
done
