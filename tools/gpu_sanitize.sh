#!/bin/bash
# compute-sanitizer passes over small kernel tests (memcheck: out-of-bounds / misaligned; racecheck: shared-memory hazards)
export PYTHONPATH=$PWD
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  echo "=== $tool"
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_ops.py -x -q \
    -k "test_attention_tcgen05_vs_oracle or test_gemm_pair_requant or test_gemm_pair_residual or test_gemm_pair_gelu or test_layernorm_int_vs_oracle or test_patchify" \
    > gpurun_out/sanitize_$tool.log 2>&1
  echo "exit $?"; grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY" gpurun_out/sanitize_$tool.log | tail -3
done
