#!/bin/bash
for r in 2960 4096 4144 5920 8192; do
  echo "resident $r"
  HELM_RESIDENT=$r python tools/sweep.py case2869pegase 32768 1 2>&1 | grep '"case"' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['cases_per_s'], d['ms'], d['steps'], d.get('fc'))"
done
