"""Run one helm_batch (for ncu captures): python tools/one_batch.py CASE B G"""
import os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, 'tests'))
from common import load_case
import helmpy_b200
name, B, G = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
case = load_case(name)
scales = np.random.default_rng(0).uniform(0.8, 1.05, B)
res = helmpy_b200.helm_batch(case, scales, mismatch=1e-8, max_coefficients=40, group=G)
print('B', B, 'G', G, 'ms', res.stats['ms_total'], 'conv', int(res.converged.sum()), 'launches', res.stats['kernel_launches'])
