#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
P2V_LIB=p2vit_b200/csrc/libp2vit_b200_trace.so python tools/pair_trace.py proj 384 > gpurun_out/trace_proj_384.log 2>&1; echo "rc $?"
P2V_LIB=p2vit_b200/csrc/libp2vit_b200_trace.so python tools/pair_trace.py qkv 384 > gpurun_out/trace_qkv_384.log 2>&1; echo "rc $?"
