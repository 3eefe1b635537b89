#!/bin/bash
# usage: tools/ab_pdl_kinds.sh <bench args>: the bench with programmatic dependent launch restricted to kernel families
# (P2VIT_PDL_KINDS mask: 1 block GEMMs, 2 attention, 4 LayerNorm, 8 the rest; 0 = none, 15 = all)
for k in 0 15 1 2 4 3 5 6; do
  P2VIT_PDL_KINDS=$k python bench.py --no-cpu-baseline "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('kinds=$k', round(d['value']), round(d['ms_per_step'],4))"
done
