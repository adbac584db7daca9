"""Exploration helper (GPU box): time helm_batch over batch sizes / group sizes and
print the per-kernel split from the library's event profile mode."""
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, 'tests'))
from common import load_case  # noqa: E402
import helmpy_b200  # noqa: E402
from helmpy_b200.core.helm import get_plan  # noqa: E402
from helmpy_b200.core import _lib  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else 'case2869pegase'
    Bs = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else '256,1024').split(',')]
    Gs = [int(x) for x in (sys.argv[3] if len(sys.argv) > 3 else '1,2,4,8,16,32').split(',')]
    profile = len(sys.argv) > 4 and sys.argv[4] == 'profile'
    resident = int(os.environ.get('HELM_RESIDENT', '0'))
    case = load_case(name)
    print('peaks', _lib.measure_peaks())
    plan = get_plan(case)
    print('plan', plan.info)
    for B in Bs:
        scales = np.random.default_rng(0).uniform(0.8, 1.05, B)
        for G in Gs:
            if profile:
                os.environ['HELM_B200_PROFILE'] = '1'
            else:
                os.environ.pop('HELM_B200_PROFILE', None)
            best = None
            for rep in range(2 if profile else 3):
                t = time.perf_counter()
                res = helmpy_b200.helm_batch(case, scales, mismatch=1e-8, max_coefficients=40, group=G, max_resident=resident)
                wall = time.perf_counter() - t
                if best is None or res.stats['ms_total'] < best[0]:
                    best = (res.stats['ms_total'], wall, res)
            ms, wall, res = best
            solves = int(np.sum(np.maximum(res.ncoef - 1, 0)))
            out = dict(case=name, B=B, G=G, ms=round(ms, 2), wall_ms=round(wall * 1e3, 1),
                       cases_per_s=round(B / ms * 1e3, 1), conv=int(res.converged.sum()),
                       steps=res.stats['steps'], scen_solves=solves,
                       fc=[res.stats.get('factor_cache_hits'), res.stats.get('factorizations')])
            if profile:
                out['ms_kernel'] = {k: round(v, 2) for k, v in res.stats['ms_kernel'].items()}
            print(json.dumps(out), flush=True)


if __name__ == '__main__':
    main()
