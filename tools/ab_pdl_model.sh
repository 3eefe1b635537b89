#!/bin/bash
# usage: tools/ab_pdl_model.sh <bench args>: the bench with P2VIT_PDL=0 and in the default mode, twice each (interleaved)
mkdir -p gpurun_out
for r in 1 2; do for m in 0 2; do
  P2VIT_PDL=$m python bench.py --no-cpu-baseline "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('PDL=$m', round(d['value']), round(d['ms_per_step'],4))"
done; done
