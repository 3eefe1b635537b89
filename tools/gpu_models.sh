#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout ${TMO:-900} "$@" > gpurun_out/$name.log 2>&1; echo "exit $?"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/$name.log').read().strip().splitlines()[-1])
    print(d['config']['workload'], '|', round(d['value'],1), 'img/s', round(d['ms_per_step'],3), 'ms | e2e', round(d['e2e']['value'],1), '|', d['roofline']['device_ms_per_step_by_family'], d['roofline'].get('gemm_tops_by_kind'))
except Exception as e:
    print('no json', e); print(open('gpurun_out/$name.log').read()[-1500:])
PY
}
run bench_vitb python bench.py --model vit_base --batch 256 --steps 10 --warmup 3 --no-cpu-baseline
run bench_vitl python bench.py --model vit_large --batch 128 --steps 5 --warmup 3 --no-cpu-baseline --bits mixed
run bench_tiny python bench.py --model deit_tiny --batch 256 --steps 20 --warmup 3 --no-cpu-baseline
run bench_swint python bench.py --model swin_tiny --batch 256 --steps 10 --warmup 3 --no-cpu-baseline
