#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
python -m pytest tests/test_gpu_ops.py -q -k "attention" > gpurun_out/ab_att_tests.log 2>&1; echo "att tests rc $?"; tail -2 gpurun_out/ab_att_tests.log
python -m pytest tests/test_gpu_model.py -q -x -k "omse or strict or non_pot" > gpurun_out/ab_att_model.log 2>&1; echo "model tests rc $?"; tail -2 gpurun_out/ab_att_model.log
python tools/att_bench.py 6; python tools/att_bench.py 12
python bench.py --model vit_base --method omse --steps 10 --warmup 3 --no-cpu-baseline --configs none --sustain 0 > gpurun_out/zp_omse.json 2>gpurun_out/zp.err || tail -3 gpurun_out/zp.err
python - <<PY
import json
d=json.load(open('gpurun_out/zp_omse.json'))
r=d['roofline']
print('vit_base omse', round(d['value']), d['ms_per_step'], r.get('device_ms_per_step_by_family'))
PY
