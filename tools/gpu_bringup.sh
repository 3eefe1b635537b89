#!/bin/bash
# first bring-up: each group in its own process with its own timeout
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; echo "=== $name" ; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?"; tail -n 25 gpurun_out/$name.log; }
run rowops python -m pytest tests/test_gpu_ops.py -q -m gpu -k "quantize or fake_quant or patchify or layernorm or softmax or minmax" -p no:cacheprovider
run gemm_simt python -m pytest tests/test_gpu_ops.py -q -m gpu -k "gemm and simt or rejects" -p no:cacheprovider
run gemm_tc python -m pytest tests/test_gpu_ops.py -q -m gpu -k "gemm_f32 and tcgen05" -p no:cacheprovider
run gemm_epi python -m pytest tests/test_gpu_ops.py -q -m gpu -k "gemm and not gemm_f32" -p no:cacheprovider
run attention python -m pytest tests/test_gpu_ops.py -q -m gpu -k "attention" -p no:cacheprovider
run model python -m pytest tests/test_gpu_model.py -q -m gpu -p no:cacheprovider
run diag python tools/diag_parity.py deit_tiny 8 8
