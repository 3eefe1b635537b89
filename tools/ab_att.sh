#!/bin/bash
# A/B of the attention kernel's SPLIT passes (P2V_ATT_SPLIT=0: 64-bit sums + reciprocal lookups): tests, then stand-alone timings
export PYTHONPATH=$PWD
mkdir -p gpurun_out
python -m pytest tests/test_gpu_ops.py -q -k "attention or softmax" > gpurun_out/ab_att_tests.log 2>&1; echo "tests(split) rc $?"; tail -2 gpurun_out/ab_att_tests.log
P2V_ATT_SPLIT=0 python -m pytest tests/test_gpu_ops.py -q -k "attention or softmax" > gpurun_out/ab_att_tests0.log 2>&1; echo "tests(nosplit) rc $?"; tail -2 gpurun_out/ab_att_tests0.log
for h in 6 12; do
  for m in 1 0; do echo "split=$m"; P2V_ATT_SPLIT=$m python tools/att_bench.py $h; done
done 2>&1 | tee gpurun_out/ab_att.log
