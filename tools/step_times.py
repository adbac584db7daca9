"""Per-step wall time from the -DHELM_TIMELINE build: python tools/step_times.py CASE B"""
import ctypes as C, os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, 'tests'))
from common import load_case
import helmpy_b200
from helmpy_b200.core import _lib
from helmpy_b200.core.helm import get_plan
name, B = sys.argv[1], int(sys.argv[2])
resident = int(sys.argv[3]) if len(sys.argv) > 3 else 0
case = load_case(name)
scales = np.random.default_rng(0).uniform(0.8, 1.05, B)
for rep in range(2):
    res = helmpy_b200.helm_batch(case, scales, mismatch=1e-8, max_coefficients=40, group=1, max_resident=resident)
print('ms', res.stats['ms_total'], 'steps', res.stats['steps'])
plan = get_plan(case)
n = 400
buf = np.zeros((n, 16, 2), dtype=np.uint64)
_lib.check(_lib.load().helm_debug_timeline(plan.handle, buf.ctypes.data_as(C.c_void_p), n), 'timeline')
names = _lib.KERNEL_NAMES
print('cycle counters (thread 0 of every solve CTA, summed):', [int(buf[0, 7 + i, 0]) for i in range(1, 8)])
rows = []
prev_end = None
for s in range(1, n):
    row = buf[s][:8]
    valid = row[:, 1] > 0
    if not valid.any():
        continue
    t0 = row[valid, 0].min(); t1 = row[valid, 1].max()
    d = {names[k]: (int(row[k, 1]) - int(row[k, 0])) / 1e3 for k in range(8) if valid[k]}
    # end of the main branch (pass_end / qlimit / advance / rhs_conv) and of the side branch (init / factor) after the step start
    d['main_end'] = max((int(row[k, 1]) - int(t0)) / 1e3 for k in (2, 3, 4, 5, 6) if valid[k]) if any(valid[k] for k in (2, 3, 4, 5, 6)) else 0
    d['side_end'] = max((int(row[k, 1]) - int(t0)) / 1e3 for k in (0, 1, 7) if valid[k]) if any(valid[k] for k in (0, 1, 7)) else 0
    d['factor_start'] = (int(row[1, 0]) - int(t0)) / 1e3 if valid[1] else 0
    rows.append((s, (int(t1) - int(t0)) / 1e3, d))
for s, dt, d in rows:
    print('%3d %7.0f us  solve %6.0f rhs %6.0f factor %6.0f asm %5.0f init %5.0f qlim %5.0f adv %4.0f pend %4.0f | main_end %5.0f side_end %5.0f fstart %5.0f' % (s, dt, d.get('solve', 0), d.get('rhs_conv', 0), d.get('factor', 0), d.get('assemble', 0), d.get('init', 0), d.get('qlimit', 0), d.get('advance', 0), d.get('pass_end', 0), d['main_end'], d['side_end'], d['factor_start']))
