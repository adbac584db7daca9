"""Triangular-solve microbenchmark (one scenario, one kernel alone): python tools/solve_time.py [case]"""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, 'tests'))
from common import load_case
from helmpy_b200.core.helm import get_plan
name = sys.argv[1] if len(sys.argv) > 1 else 'case2869pegase'
plan = get_plan(load_case(name))
for kern, kn in ((0, 'k_solve_one'), (1, 'k_solve_cta'), (2, 'k_solve_flat'), (3, 'k_solve_rows_cta'), (4, 'k_solve_rows'), (5, 'k_solve_rows_deep')):
    print(name, kn, 'us per solve: %.1f' % (1e3 * plan.debug_solve_time(kern, 0, 50)))
for abl, what in ((1, 'no shuffle rounds'), (2, 'no factor loads'), (4, 'no x gathers'), (8, 'no block barriers'), (16, 'no waves: records, barriers, prologue, epilogue'), (1 | 2 | 4, 'no shuffles/loads/gathers'), (1 | 2 | 4 | 8, 'no shuffles/loads/gathers/barriers')):
    print('  k_solve_rows_cta ablation %2d (%s): %.1f us' % (abl, what, 1e3 * plan.debug_solve_time(3, abl, 50)))
for abl, what in ((1, 'no shuffle reduction'), (2, 'no operand copies'), (4, 'no x gather'), (8, 'no block barriers'),
                  (16, 'no thin segments'), (32, 'no wide segments'), (1 | 4, 'no shuffles, no gather'),
                  (1 | 2 | 4, 'no shuffles/copies/gather'), (1 | 2 | 4 | 8, 'no shuffles/copies/gather/barriers'),
                  (16 | 32, 'no segments at all: prologue + epilogue'), (64, 'direct loads instead of the ring'), (64 | 1 | 4 | 8, 'direct loads, no shuffles/gather/barriers')):
    print('  k_solve_one ablation %2d (%s): %.1f us' % (abl, what, 1e3 * plan.debug_solve_time(0, abl, 50)))
