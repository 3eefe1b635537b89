#!/bin/bash
# does the "other" family (one-tile GEMMs, patchify, class rows, Swin window attention / gathers) gain from PDL?  mask 5 vs 13
for a in "--model swin_tiny --steps 10" "--steps 20"; do for k in 5 13; do
  P2VIT_PDL_KINDS=$k python bench.py --no-cpu-baseline $a 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$a kinds=$k', round(d['value']), round(d['ms_per_step'],4))"
done; done
