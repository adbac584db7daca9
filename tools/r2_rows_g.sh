#!/bin/bash
export HELM_RESIDENT=4096
HELM_B200_ROWS_CTA=1 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "benchmark_path or large_batch or full_size" 2>&1 | tail -2
run() {  # label, env
  echo "variant $1"
  env $2 python tools/sweep.py case2869pegase 16384 1 profile 2>&1 | grep '"case"' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms'], d['ms_kernel']['solve'], d['ms_kernel']['factor'], d['ms_kernel']['rhs_conv'], d['ms_kernel']['eps'], d['conv'])"
  env $2 python tools/sweep.py case2869pegase 16384 1 2>&1 | grep '"case"' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['cases_per_s'], d['ms'])"
}
run rows A=1
run rows_cta HELM_B200_ROWS_CTA=1
