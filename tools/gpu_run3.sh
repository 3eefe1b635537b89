#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout 1200 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?"; tail -n ${TAILN:-15} gpurun_out/$name.log; }
run ops python -m pytest tests/test_gpu_ops.py -q -m gpu -p no:cacheprovider
run model python -m pytest tests/test_gpu_model.py -q -m gpu -p no:cacheprovider
run smoke python -c "import __graft_entry__ as g; g.smoke()"
TAILN=3 run bench python bench.py --steps 20 --warmup 3
TAILN=3 run bench_golden python bench.py --steps 20 --warmup 3 --golden-state --no-cpu-baseline
TAILN=3 run bench_ref python bench.py --impl reference --steps 2 --warmup 1
echo "=== ncu launches"
timeout 900 python bench.py --steps 2 --warmup 3 --golden-state --no-cpu-baseline > gpurun_out/plain_for_ncu.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'gemm_tc|attention|layernorm|patchify|fill_cls' -c 360 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --golden-state --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
echo "exit $?"; tail -3 gpurun_out/ncu_launches.log
echo "=== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gemm_tc|attention_simt' -s 40 -c 8 -o gpurun_out/prof_r1a python bench.py --steps 2 --warmup 3 --golden-state --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
echo "exit $?"; tail -3 gpurun_out/ncu_full.log; ls -la gpurun_out/
