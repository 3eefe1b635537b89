export PYTHONPATH=$PWD
for cfg in "-1 0" "-1 128" "-1 192" "0 192" "-1 256"; do set -- $cfg; echo "== deit_small BRES=$1 BN=$2"; P2V_PAIR_BRES=$1 P2V_PAIR_BN=$2 python tools/gemm_bench.py deit_small 256 2 2>&1 | grep -E "fc2|proj"; done
for cfg in "-1 0" "-1 128" "-1 192" "-1 256" "0 256"; do set -- $cfg; echo "== vit_base BRES=$1 BN=$2"; P2V_PAIR_BRES=$1 P2V_PAIR_BN=$2 python tools/gemm_bench.py vit_base 256 2 2>&1 | grep -E "fc2|proj|fc1"; done
