"""Debug helper (GPU box): factor+solve stage on a tiled grid vs SuperLU, for the solve kernel selected by env."""
import os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, 'tests'))
from scipy.sparse.linalg import splu
import synth
from common import load_case
from oracle import helm_oracle as ho
from helmpy_b200.core.helm import get_plan
names = sys.argv[1].split(',')
case = synth.tiled_case(names) if len(names) > 1 else load_case(names[0])
plan = get_plan(case)
g = ho.Grid(case)
B = 3
bts = np.tile(g.btype0.astype(np.uint8), (B, 1))
rng = np.random.default_rng(1)
rhs = rng.standard_normal((B, 2 * case.N))
x = plan.debug_factor_solve(bts, rhs, group=1)
A, _ = ho.build_matrix(g, bts[0].astype(np.int8))
lu = splu(A)
for b in range(B):
    xs = lu.solve(rhs[b])
    err = np.abs(x[b] - xs)
    bad = np.nonzero(err > 1e-8 * np.max(np.abs(xs)))[0]
    print('scenario', b, 'max err', err.max(), 'rel', err.max() / np.max(np.abs(xs)), 'bad entries', len(bad), bad[:10] // 2)
