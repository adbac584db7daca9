#!/bin/bash
# round-2 experiment batch b: lane-per-scenario solve A/B + single-case timeline
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
export HELM_RESIDENT=4096
echo "== correctness with the lane solve"
HELM_B200_LANE_SOLVE=1 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "benchmark_path or large_batch or full_size" 2>&1 | tail -3
echo "== A/B 16384 scenarios"
for s in "" "HELM_B200_LANE_SOLVE=1"; do
  echo "setting: $s"
  env $s python tools/sweep.py case2869pegase 16384 1 2>&1 | grep '"case"'
  env $s python tools/sweep.py case2869pegase 16384 1 profile 2>&1 | grep '"case"'
done
echo "== single-case timeline"
HELM_B200_LIB=$PWD/helmpy_b200/libhelm_b200_tl.so python tools/timeline.py case2869pegase 1 1,2,3,10,11,50 2>&1 | tail -12
HELM_B200_LIB=$PWD/helmpy_b200/libhelm_b200_tl.so python tools/step_times.py case2869pegase 1 2>&1 | head -30
