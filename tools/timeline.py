"""Print the per-step kernel timeline (needs the -DHELM_TIMELINE build:
HELM_B200_LIB=helmpy_b200/libhelm_b200_tl.so python tools/timeline.py CASE B)."""
import ctypes as C, os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, 'tests'))
from common import load_case
import helmpy_b200
from helmpy_b200.core import _lib
from helmpy_b200.core.helm import get_plan
name, B = sys.argv[1], int(sys.argv[2])
steps = [int(x) for x in (sys.argv[3] if len(sys.argv) > 3 else '1,2,30,31,32,60,61').split(',')]
case = load_case(name)
scales = np.random.default_rng(0).uniform(0.8, 1.05, B)
for rep in range(2):
    res = helmpy_b200.helm_batch(case, scales, mismatch=1e-8, max_coefficients=40, group=1)
print('ms', res.stats['ms_total'], 'steps', res.stats['steps'])
plan = get_plan(case)
n = 400
buf = np.zeros((n, 16, 2), dtype=np.uint64)
_lib.check(_lib.load().helm_debug_timeline(plan.handle, buf.ctypes.data_as(C.c_void_p), n), 'timeline')
names = _lib.KERNEL_NAMES
print('solve cycle counters (thread 0 of every CTA, summed): load_rhs, ring_wait+barrier, ring_compute, wideL, wideU, store_x, between', [int(buf[0, 7 + i, 0]) for i in range(1, 8)])
for s in steps:
    row = buf[s]
    valid = row[:, 1] > 0
    if not valid.any():
        continue
    t0 = row[valid, 0].min()
    print('step', s, ' '.join('%s[%.0f..%.0f]' % (names[k], (row[k, 0] - t0) / 1e3, (row[k, 1] - t0) / 1e3)
                              for k in range(8) if valid[k]), 'us')
tot = {}
for s in range(1, n):
    row = buf[s]
    valid = row[:, 1] > 0
    if not valid.any():
        continue
    t0 = row[valid, 0].min(); t1 = row[valid, 1].max()
    tot['step'] = tot.get('step', 0) + (t1 - t0) / 1e6
    for k in range(8):
        if valid[k]:
            tot[names[k]] = tot.get(names[k], 0) + (row[k, 1] - row[k, 0]) / 1e6
print({k: round(float(v), 2) for k, v in tot.items()})
