#!/bin/bash
export HELM_RESIDENT=4096
run() {  # label, lib, env
  echo "variant $1"
  env HELM_B200_LIB=$PWD/helmpy_b200/libhelm_b200$2.so $3 python tools/sweep.py case2869pegase 16384 1 profile 2>&1 | grep '"case"' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms'], d['ms_kernel']['solve'], d['ms_kernel']['factor'], d['conv'])"
  env HELM_B200_LIB=$PWD/helmpy_b200/libhelm_b200$2.so $3 python tools/sweep.py case2869pegase 16384 1 2>&1 | grep '"case"' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['cases_per_s'], d['ms'])"
}
run f0 _f0 A=1
run f1_spill _f1 A=1
run f1_nopin _f1np A=1
run f1_r72 _f1r72 A=1
run f1_r72_nb4 _f1r72nb4 A=1
run f0_r72_nb4 _f0r72nb4 A=1
