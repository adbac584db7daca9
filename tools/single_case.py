"""Single-case latency (BASELINE config 2): python tools/single_case.py [case] [reps]"""
import os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, 'tests'))
from common import load_case, golden_results
import helmpy_b200
name = sys.argv[1] if len(sys.argv) > 1 else 'case2869pegase'
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 7
case = load_case(name)
ts = []
for _ in range(reps):
    r = helmpy_b200.helm_batch(case, [1.0], mismatch=1e-8, max_coefficients=40)
    ts.append(r.stats['ms_total'])
g = golden_results(name)
V = r.V[0]
print(name, 'single-case ms median %.3f min %.3f' % (float(np.median(ts[1:])), min(ts[1:])), 'steps', r.stats['steps'], 'coef', r.list_coef(0),
      'dmag %.2e' % float(np.max(np.abs(np.abs(V) - g['classic_mag']))), 'lib', os.environ.get('HELM_B200_LIB', 'default'))
