#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
echo "=== ln tests"
timeout 300 python -m pytest tests/test_gpu_ops.py -x -q -k "layernorm or norm" > gpurun_out/ln_tests.log 2>&1
rc=$?; echo "exit $rc"; tail -n 8 gpurun_out/ln_tests.log
[ $rc -ne 0 ] && exit 0
echo "=== model tests"
timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_swin.py -x -q > gpurun_out/model_tests.log 2>&1
echo "exit $?"; tail -n 5 gpurun_out/model_tests.log
echo "=== bench"
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1
echo "exit $?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['device_ms_per_step_by_family'], d['roofline']['gemm_ms_by_kind'])
PY
