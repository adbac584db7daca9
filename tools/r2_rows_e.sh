#!/bin/bash
export HELM_RESIDENT=4096
run() {  # label, env
  echo "variant $1"
  env $2 python tools/sweep.py case2869pegase 16384 1 profile 2>&1 | grep '"case"' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms'], d['ms_kernel']['solve'], d['ms_kernel']['factor'], d['ms_kernel']['rhs_conv'], d['ms_kernel']['eps'], d['conv'])"
  env $2 python tools/sweep.py case2869pegase 16384 1 2>&1 | grep '"case"' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['cases_per_s'], d['ms'])"
}
run default A=1
run sort_dfs HELM_B200_SORT_DFS=1
