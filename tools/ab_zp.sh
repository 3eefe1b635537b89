#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/tests_full.log 2>&1; echo "tests rc $?"; tail -3 gpurun_out/tests_full.log
for spec in "vit_base omse"; do
  set -- $spec
  python bench.py --model $1 --method $2 --steps 10 --warmup 3 --no-cpu-baseline --configs none --sustain 0 > gpurun_out/zp_$1_$2.json 2>gpurun_out/zp.err || tail -3 gpurun_out/zp.err
  python - <<PY
import json
d=json.load(open('gpurun_out/zp_$1_$2.json'))
r=d['roofline']
print('$1 $2', round(d['value']), d['ms_per_step'], r.get('device_ms_per_step_by_family'), r.get('gemm_ms_by_kind'))
PY
done
