#!/bin/bash
# mid-size batches: CTA-per-scenario deep solve (default up to 888 scenarios) against the warp-per-scenario row-lane solve
for b in 148 200 296 512 800; do
  for m in 888 100; do
    HELM_B200_CTA_SOLVE_MAX=$m python tools/sweep.py case2869pegase $b 1 2>&1 | grep '"case"' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('B', d['B'], 'cta_solve_max $m', 'ms', d['ms'], 'cases/s', d['cases_per_s'])"
  done
done
