#!/bin/bash
# long rows (13.6 k-bus grid): every solve variant against the oracle-checked wave solve
python - <<'PY'
import sys, os, time
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import numpy as np, synth
import helmpy_b200
case = synth.tiled_case(['case2869pegase'] * 4 + ['case1354pegase'] + ['case118'] * 7)
s = np.random.default_rng(9).uniform(0.9, 1.03, 1024)
out = {}
for label, env in (('rows', {}), ('flat', {'HELM_B200_ROWS_SOLVE': '0'}), ('rows_cta', {'HELM_B200_ROWS_CTA': '1'})):
    for k in ('HELM_B200_ROWS_SOLVE', 'HELM_B200_ROWS_CTA'):
        os.environ.pop(k, None)
    os.environ.update(env)
    # a fresh plan per variant: the switches are read when the plan is created
    case2 = synth.tiled_case(['case2869pegase'] * 4 + ['case1354pegase'] + ['case118'] * 7)
    r = helmpy_b200.helm_batch(case2, s, mismatch=1e-8, max_coefficients=40)
    r = helmpy_b200.helm_batch(case2, s, mismatch=1e-8, max_coefficients=40)
    out[label] = r
    print(label, 'ms', round(r.stats['ms_total'], 1), 'cases/s', round(1024 / r.stats['ms_total'] * 1e3, 1), 'conv', int(r.converged.sum()))
for k in ('HELM_B200_ROWS_SOLVE', 'HELM_B200_ROWS_CTA'):
    os.environ.pop(k, None)
for label in ('flat', 'rows_cta'):
    print(label, 'vs rows: same ncoef', bool(np.array_equal(out[label].ncoef, out['rows'].ncoef)), 'max dV', float(np.max(np.abs(out[label].V - out['rows'].V))))
small = helmpy_b200.helm_batch(case, s[:8], mismatch=1e-8, max_coefficients=40)      # k_solve_rows_deep
print('deep (8 scenarios) vs rows: same ncoef', bool(np.array_equal(small.ncoef, out['rows'].ncoef[:8])), 'max dV', float(np.max(np.abs(small.V - out['rows'].V[:8]))))
PY
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "tiled or n_minus_1 or benchmark_path or large_batch" 2>&1 | tail -2
