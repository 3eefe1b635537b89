#!/bin/bash
# 8-GPU bench line (fp32 e2e and uint8 e2e) -> gpurun_out/bench_n8.json
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err
echo "rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/bench_n8.json'))
print(round(d['value']), round(d['e2e']['value']), round(d['e2e_u8']['value']), d['ms_per_step'])"
