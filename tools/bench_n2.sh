#!/bin/bash
# 2-GPU bench line with the final build -> gpurun_out/bench_n2.json
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "rc=$? lines=$(wc -l < gpurun_out/bench_n2.json)"
python -c "
import json
d=json.load(open('gpurun_out/bench_n2.json'))
print(round(d['value']), round(d['e2e']['value']), round(d['e2e_u8']['value']), d['ms_per_step'])"
