"""How many distinct matrices does a load-scaling batch factorize?  For every Q-limit pass of every scenario
the matrix depends on the bus types only (helm.py:46-60), i.e. on the set of buses switched so far.
python tools/sharing_stats.py [case] [B]"""
import os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, 'tests'))
from common import load_case
from helmpy_b200.core.helm import get_plan
from helmpy_b200.core import _lib
name = sys.argv[1] if len(sys.argv) > 1 else 'case2869pegase'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
case = load_case(name)
plan = get_plan(case)
scales = np.random.default_rng(0).uniform(0.8, 1.05, B)
prm = _lib.Plan.make_params(1e-8, 40, True, record_switches=True)
res = plan.solve_host(scales, prm)
V, status, ncoef, nswitch = res[:4]
sw = plan.last_switches(B, nswitch)
npass = (ncoef > 0).sum(axis=1)
print(name, 'B', B, 'converged', int((status == 2).sum()), 'passes per scenario', np.bincount(npass))
total = 0
distinct_total = 0
for p in range(2, int(npass.max()) + 1):
    keys = {}
    for b in range(B):
        if npass[b] >= p:
            key = frozenset(bus for (ps, bus, q) in sw[b] if ps < p)      # switched before pass p starts
            keys[key] = keys.get(key, 0) + 1
    n = sum(keys.values())
    total += n
    distinct_total += len(keys)
    top = sorted(keys.values(), reverse=True)[:5]
    print('pass', p, 'scenarios', n, 'distinct matrices', len(keys), 'largest groups', top)
print('factorizations now', total, 'distinct', distinct_total, 'ratio %.1f' % (total / max(1, distinct_total)))
