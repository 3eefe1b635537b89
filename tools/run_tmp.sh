#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
python -m pytest tests/test_gpu_ops.py -q -k "gelu or gemm" > gpurun_out/t_ops.log 2>&1; echo "ops tests rc $?"; tail -3 gpurun_out/t_ops.log
python -m pytest tests/test_gpu_model.py tests/test_gpu_swin.py -q > gpurun_out/t_model.log 2>&1; echo "model tests rc $?"; tail -3 gpurun_out/t_model.log
python tools/gemm_bench.py deit_small 256 2 | grep -i "fc1"
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --configs none --sustain 0 > gpurun_out/g.json 2>gpurun_out/g.err || tail -3 gpurun_out/g.err
python - <<PY
import json
d=json.load(open('gpurun_out/g.json'))
r=d['roofline']
print('deit_small', round(d['value']), d['ms_per_step'], r.get('gemm_ms_by_kind'))
PY
