#!/bin/bash
mkdir -p gpurun_out
for d in 0 1 2 3; do
  P2V_DBG=$d timeout 300 python bench.py --steps 10 --warmup 3 --golden-state --no-cpu-baseline > gpurun_out/dbg$d.log 2>&1
  python -c "
import json;d=json.loads(open('gpurun_out/dbg$d.log').read().strip().splitlines()[-1]);print('dbg$d', round(d['ms_per_step'],3), d['roofline']['gemm_ms_by_kind'])"
done
