#!/bin/bash
# fast LayerNorm for non-power-of-two scales (P2V_LN_NP=0: generic kernel): parity tests, then ViT-B percentile / omse
export PYTHONPATH=$PWD
mkdir -p gpurun_out
python -m pytest tests/test_gpu_ops.py -q -k "layernorm" > gpurun_out/ab_np_tests.log 2>&1; echo "ln tests rc $?"; tail -2 gpurun_out/ab_np_tests.log
python -m pytest tests/test_gpu_model.py tests/test_gpu_swin.py -q -x > gpurun_out/ab_np_model.log 2>&1; echo "model tests rc $?"; tail -2 gpurun_out/ab_np_model.log
for spec in "vit_base percentile 1" "vit_base percentile 0" "vit_base omse 1" "deit_small ema 1" "deit_small ema 0"; do
  set -- $spec
  P2V_LN_NP=$3 python bench.py --model $1 --method $2 --steps 10 --warmup 3 --no-cpu-baseline --configs none --sustain 0 > gpurun_out/np_$1_$2_$3.json 2>gpurun_out/np.err || tail -3 gpurun_out/np.err
  python - <<PY
import json
d=json.load(open('gpurun_out/np_$1_$2_$3.json'))
r=d['roofline']
print('$1 $2 np=$3', round(d['value']), d['ms_per_step'], r.get('device_ms_per_step_by_family'))
PY
done
