"""Synthetic large grids for the configurations of BASELINE.json whose MATPOWER
cases (case2383wp, case9241pegase, case13659pegase) are not shipped by the
reference: copies of shipped grids tied together through
low-impedance lines between their former slack buses.  Extra slacks become
generator (PV) buses carrying the active power the slack produced in the
stand-alone solution, loads are perturbed per copy so the copies differ."""
import numpy as np

from common import load_tables, load_case


def _slack_generation(name):
    from oracle import helm_oracle as ho
    case = load_case(name)
    V, _ = ho.helm_oracle(case, mismatch=1e-8, max_coefficients=40)
    rows = np.repeat(np.arange(case.N), np.diff(case.adj_ptr))
    I = np.zeros(case.N, dtype=complex)
    np.add.at(I, rows, case.Y_val * V[case.adj_idx])
    S = V * np.conj(I)
    s = case.slack
    return (S[s].real + case.Pd[s]) * 100.0      # MW


def tile_tables(names, seed=0, load_jitter=0.02, tie_x=0.01, star=True):
    """Return (buses, branches, generators) of the tiled grid."""
    rng = np.random.default_rng(seed)
    B, BR, G = [], [], []
    offset = 0
    slack_ids = []
    for k, name in enumerate(names):
        b, br, g = [a.copy() for a in load_tables(name)]
        ids = b[:, 0].copy()
        remap = {int(v): offset + i + 1 for i, v in enumerate(ids)}
        b[:, 0] = [remap[int(v)] for v in ids]
        br[:, 0] = [remap[int(v)] for v in br[:, 0]]
        br[:, 1] = [remap[int(v)] for v in br[:, 1]]
        g[:, 0] = [remap[int(v)] for v in g[:, 0]]
        jitter = 1.0 + load_jitter * rng.uniform(-1, 1, b.shape[0])
        b[:, 2] *= jitter
        b[:, 3] *= jitter
        srow = int(np.nonzero(b[:, 1] == 3)[0][0])
        slack_ids.append(int(b[srow, 0]))
        if k > 0:                                   # demote: PV bus with the slack's stand-alone output
            b[srow, 1] = 2
            grow = int(np.nonzero(g[:, 0] == b[srow, 0])[0][0])
            g[grow, 1] = _slack_generation(name)
        B.append(b)
        BR.append(br)
        G.append(g)
        offset += b.shape[0]
    buses = np.vstack(B)
    gens = np.vstack(G)
    ncol = BR[0].shape[1]
    ties = []
    for k in range(1, len(names)):
        t = np.zeros(ncol)
        t[0], t[1] = slack_ids[0 if star else k - 1], slack_ids[k]
        t[2], t[3], t[4] = 0.0, tie_x, 0.0
        t[8], t[9] = 0.0, 0.0
        if ncol > 10:
            t[10] = 1
        ties.append(t)
    branches = np.vstack(BR + ([np.array(ties)] if ties else []))
    return buses, branches, gens


def tiled_case(names, seed=0, name=None):
    from helmpy_b200.core.classes import create_case_data_object_from_arrays
    b, br, g = tile_tables(names, seed=seed)
    return create_case_data_object_from_arrays(b, br, g, name or ('tiled_' + '_'.join(names)))
