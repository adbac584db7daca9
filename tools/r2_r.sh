#!/bin/bash
python tools/single_case.py
HELM_B200_NO_FACTOR_CTA=1 python tools/single_case.py
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/sweep.py case2869pegase 64,128 1 2>&1 | grep '"case"'
HELM_B200_NO_FACTOR_CTA=1 python tools/sweep.py case2869pegase 64,128 1 2>&1 | grep '"case"'
