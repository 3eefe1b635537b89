#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
python -m pytest tests/test_gpu_ops.py -q -k "gelu or gemm" > gpurun_out/t_ops.log 2>&1; echo "ops tests rc $?"; tail -3 gpurun_out/t_ops.log
python -m pytest tests/test_gpu_model.py -q > gpurun_out/t_model.log 2>&1; echo "model tests rc $?"; tail -3 gpurun_out/t_model.log
python bench.py --model vit_base --method omse --steps 10 --warmup 3 --no-cpu-baseline --configs none --sustain 0 > gpurun_out/zp_omse.json 2>gpurun_out/zp.err || tail -3 gpurun_out/zp.err
python - <<PY
import json
d=json.load(open('gpurun_out/zp_omse.json'))
r=d['roofline']
print('vit_base omse', round(d['value']), d['ms_per_step'], r.get('device_ms_per_step_by_family'), r.get('gemm_ms_by_kind'))
PY
