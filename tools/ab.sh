#!/bin/bash
# A/B of one environment switch on one command, same box, same call:  tools/ab.sh VAR "v1 v2 ..." <command ...>
#   tools/ab.sh P2V_GELU_GUARD "0 1" python tools/gemm_bench.py deit_small 256 2
#   tools/ab.sh P2V_ATT_EXACT "0 1" python bench.py --model vit_base --method percentile --steps 10 --no-cpu-baseline --configs none --sustain 0
# (the switches are listed in INTEGRATION.md; every one of them leaves the results bit-identical)
export PYTHONPATH=$PWD
var=$1; vals=$2; shift 2
for v in $vals; do
  echo "== $var=$v"
  env $var=$v "$@" 2>&1 | tail -n 3 | cut -c1-600
done
