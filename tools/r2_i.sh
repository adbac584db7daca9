#!/bin/bash
export HELM_RESIDENT=4096
export HELM_B200_LANE_SOLVE=1
HELM_B200_LANE_SOLVE=1 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "benchmark_path or large_batch" 2>&1 | tail -2
for v in "" _ln_16_4_1 _ln_32_2_1; do
  echo "variant $v"
  HELM_B200_LIB=$PWD/helmpy_b200/libhelm_b200$v.so python tools/sweep.py case2869pegase 16384 1 profile 2>&1 | grep '"case"' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms'], d['ms_kernel']['solve'], d['conv'])"
done
echo "idx global"
HELM_B200_LANE_IDX_GLOBAL=1 python tools/sweep.py case2869pegase 16384 1 profile 2>&1 | grep '"case"' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms'], d['ms_kernel']['solve'], d['conv'])"
HELM_B200_LIB=$PWD/helmpy_b200/libhelm_b200_tl.so python tools/timeline.py case2869pegase 4096 30,31 2>&1 | grep -v "^step"
