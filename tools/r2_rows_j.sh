#!/bin/bash
python tools/solve_time.py 2>&1 | head -6
python tools/single_case.py
HELM_B200_ROWS_DEEP=1 python tools/single_case.py
HELM_B200_ROWS_DEEP=1 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "golden_xlsx or reference_runs or oracle_random or graph_and_eager or per_order" 2>&1 | tail -2
