#!/bin/bash
# A/B of the clean-table GELU epilogue (P2V_GELU_GUARD=1 keeps the near-threshold distance test)
export PYTHONPATH=$PWD
mkdir -p gpurun_out
python -m pytest tests/test_gpu_ops.py -q -k "gelu or gemm" > gpurun_out/ab_gelu_tests.log 2>&1; echo "tests rc $?"; tail -2 gpurun_out/ab_gelu_tests.log
for m in 0 1 0 1; do echo "guard=$m"; P2V_GELU_GUARD=$m python tools/gemm_bench.py deit_small 256 2 | grep -i fc1; done 2>&1 | tee gpurun_out/ab_gelu.log
for m in 0 1; do echo "guard=$m"; P2V_GELU_GUARD=$m python tools/gemm_bench.py vit_base 256 2 | grep -i fc1; done 2>&1 | tee -a gpurun_out/ab_gelu.log
for m in 0 1; do
P2V_GELU_GUARD=$m python bench.py --steps 20 --warmup 3 --no-cpu-baseline --configs none --sustain 0 --golden-state > gpurun_out/ab_bench$m.json 2>gpurun_out/ab_bench.err; echo "bench guard=$m rc $?"
python - <<PY
import json
d=json.load(open('gpurun_out/ab_bench$m.json'))
print(d['value'], d['ms_per_step'], d['roofline']['gemm_ms_by_kind'])
PY
done
