#!/bin/bash
export HELM_RESIDENT=4096
echo "== correctness with the lane solve"
HELM_B200_LANE_SOLVE=1 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "benchmark_path or large_batch or full_size or factor_solve" 2>&1 | tail -3
echo "== A/B 16384 scenarios"
for s in "" "HELM_B200_LANE_SOLVE=1"; do
  echo "setting: $s"
  env $s python tools/sweep.py case2869pegase 16384 1 2>&1 | grep '"case"'
  env $s python tools/sweep.py case2869pegase 16384 1 profile 2>&1 | grep '"case"'
done
