export PYTHONPATH=$PWD
for m in vit_base vit_large; do
for cfg in "-1 0" "-1 128" "-1 192" "-1 256" "0 192" "0 256"; do set -- $cfg; echo "== $m BRES=$1 BN=$2"; P2V_PAIR_BRES=$1 P2V_PAIR_BN=$2 python tools/gemm_bench.py $m 128 2 2>&1 | grep -v identical; done; done
