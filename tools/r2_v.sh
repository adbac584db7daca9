#!/bin/bash
export HELM_RESIDENT=4096
for v in _old "" _fl_9_56 _fl_10_48 _fl_8_64; do
  echo "variant $v"
  HELM_B200_LIB=$PWD/helmpy_b200/libhelm_b200$v.so python tools/sweep.py case2869pegase 16384 1 profile 2>&1 | grep '"case"' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms'], d['ms_kernel']['solve'], d['conv'])"
  HELM_B200_LIB=$PWD/helmpy_b200/libhelm_b200$v.so python tools/sweep.py case2869pegase 16384 1 2>&1 | grep '"case"' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['cases_per_s'], d['ms'])"
done
