#!/bin/bash
export HELM_RESIDENT=4096
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -k "factor_cache or benchmark_path or large_batch or full_size or continuous or n_minus_1 or tiled or islanding or failed_outage" 2>&1 | tail -2
run() {  # label, lib, env
  echo "variant $1"
  env HELM_B200_LIB=$PWD/helmpy_b200/libhelm_b200$2.so $3 python tools/sweep.py case2869pegase 16384 1 profile 2>&1 | grep '"case"' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms'], d['ms_kernel'], d['conv'])"
  env HELM_B200_LIB=$PWD/helmpy_b200/libhelm_b200$2.so $3 python tools/sweep.py case2869pegase 16384 1 2>&1 | grep '"case"' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['cases_per_s'], d['ms'], d['steps'], d.get('fc'))"
}
run cache "" A=1
