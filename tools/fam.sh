#!/bin/bash
# per-family device times of the other BASELINE configs (the default bench line carries them for DeiT-S only)
export PYTHONPATH=$PWD
mkdir -p gpurun_out
for spec in "vit_base minmax" "vit_base percentile" "vit_base omse" "swin_tiny minmax" "deit_tiny minmax"; do
  set -- $spec
  python bench.py --model $1 --method $2 --steps 10 --warmup 3 --no-cpu-baseline --configs none --sustain 0 > gpurun_out/fam_$1_$2.json 2>gpurun_out/fam.err || tail -3 gpurun_out/fam.err
  python - <<PY
import json
d=json.load(open('gpurun_out/fam_$1_$2.json'))
r=d['roofline']
print('$1 $2', round(d['value']), d['ms_per_step'], r.get('device_ms_per_step_by_family'), r.get('gemm_ms_by_kind'))
PY
done
