#!/bin/bash
export HELM_RESIDENT=4096
python tools/single_case.py
HELM_B200_ONE_SOLVE=0 python tools/single_case.py
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
echo "== batch: one-solve for everything"
HELM_B200_ONE_SOLVE=2 python tools/sweep.py case2869pegase 16384 1 profile 2>&1 | grep '"case"' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms'], d['ms_kernel']['solve'], d['conv'])"
HELM_B200_ONE_SOLVE=2 python tools/sweep.py case2869pegase 16384 1 2>&1 | grep '"case"'
python tools/sweep.py case2869pegase 512 1 2>&1 | grep '"case"'
HELM_B200_ONE_SOLVE=0 python tools/sweep.py case2869pegase 512 1 2>&1 | grep '"case"'
