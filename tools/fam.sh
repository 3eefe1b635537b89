#!/bin/bash
# per-family device times of the other BASELINE configs (the default bench line carries them for DeiT-S only) -> profiles/r2_family_times.txt
export PYTHONPATH=$PWD
mkdir -p gpurun_out
: > gpurun_out/family_times.txt
for spec in "deit_small minmax 8 256" "deit_tiny minmax 8 256" "vit_base minmax 8 256" "vit_base percentile 8 256" "vit_base omse 8 256" "deit_small ema 8 256" "swin_tiny minmax 8 256" "vit_large minmax mixed 1024"; do
  set -- $spec
  python bench.py --model $1 --method $2 --bits $3 --batch $4 --steps 10 --warmup 3 --no-cpu-baseline --configs none --sustain 0 > gpurun_out/fam_$1_$2.json 2>gpurun_out/fam.err || tail -3 gpurun_out/fam.err
  python - <<PY >> gpurun_out/family_times.txt
import json
d=json.load(open('gpurun_out/fam_$1_$2.json'))
r=d['roofline']
print('$1 $2 bits=$3 batch=$4: %d img/s, %.3f ms/step; families (eager, ms/step): %s; GEMMs: %s; tensor fraction of the forward %.3f' % (round(d['value']), d['ms_per_step'], r.get('device_ms_per_step_by_family'), r.get('gemm_ms_by_kind'), r.get('tensor_fraction_of_whole_forward') or 0))
PY
done
cat gpurun_out/family_times.txt
