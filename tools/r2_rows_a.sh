#!/bin/bash
# A/B of the row-lane solve (k_solve_rows) against the wave solve (k_solve_flat)
export HELM_RESIDENT=4096
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "benchmark_path or large_batch or full_size or tiled" 2>&1 | tail -3
for s in "HELM_B200_ROWS_SOLVE=0" "HELM_B200_ROWS_SOLVE=1"; do
  echo "setting: $s"
  env $s python tools/sweep.py case2869pegase 16384 1 profile 2>&1 | grep '"case"' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms'], d['ms_kernel'], d['conv'])"
  env $s python tools/sweep.py case2869pegase 16384 1 2>&1 | grep '"case"' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['cases_per_s'], d['ms'])"
done
