#!/bin/bash
export HELM_RESIDENT=4096
run() {  # label, lib, env
  echo "variant $1"
  env HELM_B200_LIB=$PWD/helmpy_b200/libhelm_b200$2.so $3 python tools/sweep.py case2869pegase 16384 1 profile 2>&1 | grep '"case"' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms'], d['ms_kernel']['solve'], d['ms_kernel']['factor'], d['ms_kernel']['fcache'], d['conv'])"
  env HELM_B200_LIB=$PWD/helmpy_b200/libhelm_b200$2.so $3 python tools/sweep.py case2869pegase 16384 1 2>&1 | grep '"case"' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['cases_per_s'], d['ms'], d.get('fc'))"
}
run cache "" A=1
run cache_nostream _nostream A=1
run cache_depth1 "" HELM_B200_SIDE_DEPTH=1
run cache_flat "" HELM_B200_ROWS_SOLVE=0
run cache_rowscta "" HELM_B200_ROWS_CTA=1
