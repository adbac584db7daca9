#!/bin/bash
export HELM_RESIDENT=4096
export HELM_B200_LANE_SOLVE=1
for v in 16_4_0 16_4_1 32_2_0 32_2_1 32_4_1; do
  echo "variant $v"
  HELM_B200_LIB=$PWD/helmpy_b200/libhelm_b200_ln_$v.so python tools/sweep.py case2869pegase 16384 1 profile 2>&1 | grep '"case"' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms'], d['ms_kernel']['solve'], d['conv'])"
done
