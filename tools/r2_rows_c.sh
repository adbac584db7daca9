#!/bin/bash
# register / blocks-per-lane variants of k_solve_rows
export HELM_RESIDENT=4096
run() {  # label, lib suffix, extra env
  echo "variant $1"
  env HELM_B200_LIB=$PWD/helmpy_b200/libhelm_b200$2.so $3 python tools/sweep.py case2869pegase 16384 1 profile 2>&1 | grep '"case"' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms'], d['ms_kernel']['solve'], d['ms_kernel']['factor'], d['conv'])"
  env HELM_B200_LIB=$PWD/helmpy_b200/libhelm_b200$2.so $3 python tools/sweep.py case2869pegase 16384 1 2>&1 | grep '"case"' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['cases_per_s'], d['ms'])"
}
run flat "" HELM_B200_ROWS_SOLVE=0
run r72 "" A=1
run r64 _r64 A=1
run r64nb3 _r64nb3 HELM_B200_ROWS_T=3
run r64nb3u2pin _r64nb3u2 HELM_B200_ROWS_T=3
run r72u2 _r72u2 A=1
run r64nb3_T2 _r64nb3 HELM_B200_ROWS_T=2
