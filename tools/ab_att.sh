#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/tests_full.log 2>&1; echo "all gpu tests rc $?"; tail -4 gpurun_out/tests_full.log
