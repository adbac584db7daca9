#!/bin/bash
export HELM_RESIDENT=4096
for v in _noearly ""; do
  echo "variant $v"
  HELM_B200_LIB=$PWD/helmpy_b200/libhelm_b200$v.so python tools/sweep.py case2869pegase 16384 1 profile 2>&1 | grep '"case"' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms'], d['ms_kernel']['rhs_conv'], d['ms_kernel']['eps'], d['conv'])"
  HELM_B200_LIB=$PWD/helmpy_b200/libhelm_b200$v.so python tools/sweep.py case2869pegase 16384 1 2>&1 | grep '"case"' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['cases_per_s'], d['ms'])"
done
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "benchmark_path or golden_xlsx or per_order" 2>&1 | tail -1
