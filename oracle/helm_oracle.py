"""TEST INFRASTRUCTURE — CPU oracle for the HELM + Pade hot path.

A numpy/scipy restatement of the reference's algorithm (HELMpy,
helmpy/core/helm.py and analytic_continuation.py), vectorised over buses but
with every floating-point sum taken in the reference's order, so its voltages
agree with the reference to round-off.  It is the checker for the CUDA path:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it.  The product (``helmpy_b200``) never
does, and has no CPU fallback.

Pinned against (tests/test_oracle.py):
  * the reference's golden result files ``data/results/Results HELM <case> 1
    1e-08.xlsx`` (copied as compact fixtures to tests/golden/ by
    oracle/gen_golden.py), and
  * outputs of the unmodified reference run in the build container
    (oracle/gen_golden.py -> tests/golden/ref_*.npz), including off-golden
    load scales, per-pass coefficient counts and switched buses.

Third-party arithmetic on the path, as in the reference: SuperLU through
``scipy.sparse.linalg.factorized`` (helm.py:106) and LAPACK zgesv through
``numpy.linalg.solve`` (analytic_continuation.py:39); neither is pinned by the
reference (setup.py:27-28).

Scope: every variant of the reference's own test matrix (test/test_helmpy.py:111-122):
pv_bus_model 1 / 2, classic slack, distributed-slack methods 1 / 2 with or without
participation factors.
"""

import cmath

import numpy as np
from scipy.sparse import csc_matrix, coo_matrix
from scipy.sparse.linalg import factorized

PQ, PV, SLACK = 0, 1, 2


def pade(series, m):
    """[L/L] Pade approximant at s=1 from the first m terms
    (analytic_continuation.py:29-50)."""
    s = np.asarray(series[:m], dtype=complex)
    L = (len(s) - 1) // 2
    C = np.zeros((L, L), dtype=complex)
    for r in range(L):
        C[r, :] = s[r + 1:r + 1 + L]                 # :35-38  C[r][c] = s[r+c+1]
    vec_b = -np.linalg.solve(C, s[L + 1:len(s)])     # :39
    b = np.ones(L + 1, dtype=complex)
    b[1:] = vec_b[::-1]                              # :40-42
    a = np.zeros(L + 1, dtype=complex)
    a[0] = s[0]
    for i in range(1, L + 1):                        # :43-49
        aux = 0
        for k in range(i + 1):
            aux += s[k] * b[i - k]
        a[i] = aux
    return np.sum(a) / np.sum(b)                     # :50


def pade_batch(V, m):
    """``pade`` for every row of V[:, :m] at once (same LAPACK routine per row)."""
    s = np.asarray(V[:, :m], dtype=complex)
    nb = s.shape[0]
    L = (m - 1) // 2
    C = np.empty((nb, L, L), dtype=complex)
    for r in range(L):
        C[:, r, :] = s[:, r + 1:r + 1 + L]
    vec_b = -np.linalg.solve(C, s[:, L + 1:m][:, :, None])[:, :, 0]
    b = np.ones((nb, L + 1), dtype=complex)
    b[:, 1:] = vec_b[:, ::-1]
    a = np.zeros((nb, L + 1), dtype=complex)
    a[:, 0] = s[:, 0]
    for i in range(1, L + 1):
        aux = np.zeros(nb, dtype=complex)
        for k in range(i + 1):
            aux = aux + s[:, k] * b[:, i - k]
        a[:, i] = aux
    return np.sum(a, axis=1) / np.sum(b, axis=1)


def epsilon(series, m):
    """Wynn epsilon (analytic_continuation.py:9-26; defined there, unused)."""
    s = np.asarray(series[:m], dtype=complex)
    n = len(s)
    e = np.zeros((n, n + 1), dtype=complex)
    e[0, 1] = s[0]
    for i in range(1, n):
        e[i, 1] = e[i - 1, 1] + s[i]
    for col in range(2, n + 1):
        for row in range(0, n + 1 - col):
            e[row, col] = e[row + 1, col - 2] + 1 / (e[row + 1, col - 1] - e[row, col - 1])
    return e[0, n]


class Grid:
    """Flat arrays the recursion needs, taken from a CaseData-like object
    (reference classes.py:180-225 or helmpy_b200.core.classes.CaseData)."""

    def __init__(self, case):
        N = self.N = case.N
        self.slack = int(case.slack)
        adj = case.branches_buses
        self.ptr = np.zeros(N + 1, dtype=np.int64)
        for i in range(N):
            self.ptr[i + 1] = self.ptr[i] + len(adj[i])
        self.idx = np.concatenate([np.asarray(r, dtype=np.int64) for r in adj])
        self.row = np.repeat(np.arange(N), np.diff(self.ptr))
        if getattr(case, 'Ytrans_val', None) is not None:
            self.ytr = np.asarray(case.Ytrans_val)
            self.yfull = np.asarray(case.Y_val)
        else:
            self.ytr = np.asarray(case.Ytrans)[self.row, self.idx]
            self.yfull = np.asarray(case.Y)[self.row, self.idx]
        self.deg = np.diff(self.ptr)
        self.Yshunt = np.array(case.Yshunt, dtype=complex)
        self.conduc = np.array(case.conduc_buses, dtype=bool)
        self.phase_bus = np.array(case.phase_barras, dtype=bool)
        self.phase = {int(k): (list(v[0]), list(v[1])) for k, v in case.phase_dict.items()}
        self.Vsp = np.array(case.V, dtype=float)
        self.Pd = np.array(case.Pd, dtype=float)
        self.Qd = np.array(case.Qd, dtype=float)
        self.Pg = np.array(case.Pg, dtype=float)
        self.Qgmax = np.array(case.Qgmax, dtype=float)
        self.Qgmin = np.array(case.Qgmin, dtype=float)
        self.list_gen = np.array(case.list_gen, dtype=np.int64)
        bt = np.zeros(N, dtype=np.int8)
        for i, t in enumerate(case.Buses_type):
            bt[i] = SLACK if t == 'Slack' else (PQ if t == 'PQ' else PV)
        self.btype0 = bt


def build_matrix(g, btype, pv_bus_model=2, DSB_model_method=None, K=None, list_gen=None):
    """Real recursion matrix (helm.py:30-107), CSC without explicit zeros.

    Returns (A, pv1_columns) where pv1_columns[i] = dict row -> removed entry of
    column 2i (pv_bus_model 1, helm.py:72-103)."""
    N = g.N
    length = 2 * N if DSB_model_method is None else 2 * N + 1
    rt = btype[g.row]
    G = g.ytr.real
    B = g.ytr.imag
    r2, c2 = 2 * g.row, 2 * g.idx
    ns = rt != SLACK
    full = (rt == PQ) | ((rt == PV) & (pv_bus_model == 1))        # helm.py:50
    rows = [r2[ns], r2[ns], r2[full] + 1, r2[full] + 1]
    cols = [c2[ns], c2[ns] + 1, c2[full], c2[full] + 1]
    vals = [G[ns], -B[ns], B[full], G[full]]
    sl = np.nonzero(btype == SLACK)[0]
    pv = np.nonzero(btype == PV)[0]
    rows += [2 * sl, 2 * sl + 1]
    cols += [2 * sl, 2 * sl + 1]
    vals += [np.ones(len(sl)), np.ones(len(sl))]
    if pv_bus_model == 2:
        rows.append(2 * pv + 1)
        cols.append(2 * pv)
        vals.append(np.ones(len(pv)))                             # helm.py:57
    if DSB_model_method is not None:                              # helm.py:62-70
        s = g.slack
        p0, p1 = g.ptr[s], g.ptr[s + 1]
        rows += [np.full(p1 - p0, 2 * N), np.full(p1 - p0, 2 * N)]
        cols += [2 * g.idx[p0:p1], 2 * g.idx[p0:p1] + 1]
        vals += [G[p0:p1], -B[p0:p1]]
        lg = np.asarray(list_gen, dtype=np.int64)
        rows += [2 * lg, np.array([2 * N])]
        cols += [np.full(len(lg), 2 * N), np.array([2 * N])]
        vals += [-K[lg], np.array([-K[s]])]
    rows, cols, vals = map(np.concatenate, (rows, cols, vals))
    A = coo_matrix((vals, (rows, cols)), shape=(length, length)).tolil()
    pv1_columns = {}
    if pv_bus_model == 1:                                         # helm.py:72-103
        for i in list_gen:
            col = {}
            for k in g.idx[g.ptr[i]:g.ptr[i + 1]]:
                col[2 * k] = A[2 * k, 2 * i]
                col[2 * k + 1] = A[2 * k + 1, 2 * i]
                A[2 * k, 2 * i] = 0
                A[2 * k + 1, 2 * i] = 0
            if DSB_model_method is not None and g.slack in g.idx[g.ptr[i]:g.ptr[i + 1]]:
                col[2 * N] = A[2 * N, 2 * i]
                A[2 * N, 2 * i] = 0
            pv1_columns[int(i)] = col
            A[2 * i + 1, 2 * i] = 1
    A = A.tocsc()
    A.eliminate_zeros()        # csc_matrix(dense) stores no explicit zeros (helm.py:106)
    A.sort_indices()
    return A, pv1_columns


def _spmv_rows(g, rows, vec):
    """sum_k Ytrans[i,k] * vec[k] over adj(i) in ascending k, for i in rows
    (sequential order of helm.py:310-312)."""
    out = np.zeros(len(rows), dtype=complex)
    start = g.ptr[rows]
    deg = g.deg[rows]
    for p in range(int(deg.max()) if len(rows) else 0):
        m = deg > p
        pos = start[m] + p
        out[m] = out[m] + g.ytr[pos] * vec[g.idx[pos]]
    return out


def helm_oracle(case, mismatch=1e-4, scale=1, max_coefficients=100, enforce_Q_limits=True,
                record=False, pv_bus_model=2, DSB_model=False, DSB_model_method=None):
    """Restatement of ``helm(case, ...)`` (helm.py:896-985) for every variant:
    pv_bus_model 1 / 2, classic slack, distributed slack methods 1 / 2.

    Does not mutate ``case``.  Returns (V complex128[N] or None, info) where
    info = {'list_coef': [...], 'switched': [...], 'passes': [per-pass dict], 'Ploss': ...}.
    """
    if DSB_model and DSB_model_method is None:
        DSB_model_method = 2                                      # helm.py:913-914
    g = Grid(case)
    N = g.N
    dsb = DSB_model_method is not None
    length = 2 * N + 1 if dsb else 2 * N
    # set_scale (classes.py:215-219): three in-place multiplications
    Pd = g.Pd * scale if scale != 1 else g.Pd.copy()
    Qd = g.Qd * scale if scale != 1 else g.Qd.copy()
    Pg = g.Pg * scale if scale != 1 else g.Pg.copy()
    Pg_imbalance = np.sum(Pd) - np.sum(Pg)                        # classes.py:286
    btype = g.btype0.copy()
    Qg = np.zeros(N)
    K = np.zeros(N)
    list_gen = g.list_gen.copy()
    list_coef, switched, passes = [], [], []
    max_coef = max_coefficients
    Vprof = np.empty(N, dtype=complex)
    phase_items = [(i, np.asarray(v[0], dtype=np.int64), np.asarray(v[1], dtype=complex))
                   for i, v in sorted(g.phase.items())]
    phase_map = {i: (nb, val) for (i, nb, val) in phase_items}
    s_ = g.slack
    m = 1

    def phase_pp(i, vecs):
        """sum_k phase_ik * vecs[k] (helm.py:375-377 and alike)"""
        if i not in phase_map:
            return 0
        nb, val = phase_map[i]
        acc = 0
        for k in range(len(nb)):
            acc += val[k] * vecs[nb[k]]
        return acc

    while True:                                                   # helm.py:936
        if DSB_model:                                             # compute_k_factor, helm.py:493-516
            K = np.zeros(N)
            Pg[s_] = Pg_imbalance
            contrib = [int(i) for i in list_gen if Pg[i] > 0]
            tot = sum(Pg[i] for i in contrib)
            if Pg[s_] > 0:
                tot += Pg[s_]
                contrib.append(s_)
            for i in contrib:
                K[i] = Pg[i] / tot
        elif dsb:                                                 # K_slack_1, helm.py:519-525
            K = np.zeros(N)
            K[s_] = 1
        A, pv1_cols = build_matrix(g, btype, pv_bus_model, DSB_model_method, K, list_gen)
        solve = factorized(A)                                     # helm.py:106
        pq = np.nonzero(btype == PQ)[0]
        pv = np.nonzero(btype == PV)[0]
        V = np.zeros((N, max_coef + 1), dtype=complex)
        W = np.zeros((N, max_coef + 1), dtype=complex)
        CC = np.zeros((len(pv), max_coef + 1), dtype=complex)     # barras_CC
        slack_CC = np.zeros(max_coef + 1, dtype=complex)
        Vre_PV = np.zeros((N, max_coef + 1))
        coeff = np.zeros((length, max_coef + 1))
        rhs_hist = np.zeros((length, max_coef + 1))
        VVant = np.zeros(len(pv))
        V[:, 0] = 1                                               # helm.py:118-135, 551-556
        W[:, 0] = 1
        coeff[0:2 * N:2, 0] = 1
        if pv_bus_model == 1:
            coeff[2 * pv, 0] = 0                                  # helm.py:124-126
            Vre_PV[:, 0] = 1
        Pi = Pg - Pd                                              # helm.py:559
        Si = Pi + Qg * 1j - Qd * 1j                               # helm.py:560
        pv_pos = {int(b): k for k, b in enumerate(pv)}
        n = 0
        m = 1
        diverged = False
        recalc = False
        while True:                                               # helm.py:567
            n += 1
            rhs = np.zeros(length)
            if pv_bus_model == 1:                                 # Calculo_Vre_PV, helm.py:142-159
                for i in list_gen:
                    if n > 1:
                        aux = 0
                        for k in range(1, n):
                            aux += V[i, k] * np.conj(V[i, n - k])
                        Vre_PV[i, n] = (-aux / 2).real
                    else:
                        Vre_PV[i, n] = (g.Vsp[i] ** 2 - 1) / 2
            # ---- PQ buses (helm.py:367-385)
            PPq = np.zeros(N, dtype=complex)
            for (i, nb, val) in phase_items:
                PPq[i] = phase_pp(i, V[:, n - 1])
            res = np.conj(Si[pq]) * np.conj(W[pq, n - 1]) - g.Yshunt[pq] * V[pq, n - 1] - PPq[pq]
            rhs[2 * pq] = res.real
            rhs[2 * pq + 1] = res.imag
            # ---- slack (helm.py:388-398)
            if n == 1:
                rhs[2 * s_] = g.Vsp[s_] - 1
            # ---- PV buses
            if len(pv) and pv_bus_model == 2:                     # helm.py:293-364
                if n == 1:
                    cc = Pi[pv] - g.Yshunt[pv].real
                    for (i, nb, val) in phase_items:
                        if i in pv_pos:
                            for v_ in val:
                                cc[pv_pos[i]] -= v_.real
                    vv = g.Vsp[pv] ** 2 - 1
                else:
                    PP = _spmv_rows(g, pv, V[:, n - 1])
                    CC[:, n - 1] = PP
                    PPP = np.zeros(len(pv), dtype=complex)
                    for x in range(1, n - 1):
                        PPP = PPP + np.conj(V[pv, n - x]) * CC[:, x]
                    PPP = PPP + np.conj(V[pv, 1]) * PP
                    cc = 0 - PPP.real
                    cd = g.conduc[pv]
                    if n > 2:
                        cc[cd] -= g.Yshunt[pv][cd].real * (VVant[cd] + 2 * V[pv, n - 1][cd].real)
                    else:
                        cc[cd] -= g.Yshunt[pv][cd].real * 2 * V[pv, 1][cd].real
                    for (i, nb, val) in phase_items:
                        if i in pv_pos:
                            acc2 = 0
                            for x in range(n):
                                acc2 += np.conj(V[i, x]) * phase_pp(i, V[:, n - 1 - x])
                            cc[pv_pos[i]] -= acc2.real
                    VVc = np.zeros(len(pv), dtype=complex)
                    for k in range(1, n):
                        VVc = VVc + V[pv, k] * np.conj(V[pv, n - k])
                    VVant = VVc.real.copy()          # stored into a float64 array (helm.py:360)
                    vv = -VVc.real
                rhs[2 * pv] = cc
                rhs[2 * pv + 1] = vv / 2
            elif len(pv):                                         # pv model 1, helm.py:261-290
                aux = np.zeros(len(pv), dtype=complex)
                auxP = np.zeros(len(pv), dtype=complex)
                for k in range(1, n):
                    aux = aux + coeff[2 * pv, k] * np.conj(W[pv, n - k])
                    if dsb:
                        auxP = auxP + coeff[2 * N, k] * np.conj(W[pv, n - k])
                res = Pi[pv] * np.conj(W[pv, n - 1]) - g.Yshunt[pv] * V[pv, n - 1] - PPq[pv] \
                    - aux * 1j + K[pv] * auxP
                rhs[2 * pv] = res.real
                rhs[2 * pv + 1] = res.imag
            # ---- distributed slack: extra equation (helm.py:164-258)
            if DSB_model_method == 1:
                auxP = 0
                for k in range(1, n):
                    auxP += coeff[2 * N, k] * np.conj(W[s_, n - k])
                res = Pi[s_] * np.conj(W[s_, n - 1]) - g.Yshunt[s_] * V[s_, n - 1] \
                    - phase_pp(s_, V[:, n - 1]) + K[s_] * auxP
                rhs[2 * N] = np.real(res)
            elif DSB_model_method == 2:
                if n > 2:
                    cc_s = 0
                    PPP = 0
                    for x in range(1, n - 1):
                        PPP += np.conj(V[s_, n - x]) * slack_CC[x]
                    PP = _spmv_rows(g, np.array([s_]), V[:, n - 1])[0]
                    slack_CC[n - 1] = PP
                    PPP += np.conj(V[s_, 1]) * PP
                    cc_s -= PPP.real
                    if g.conduc[s_]:
                        VV = 0
                        for k in range(1, n - 1):
                            VV += V[s_, k] * np.conj(V[s_, n - k])            # helm.py:219-222 (as written)
                        cc_s -= np.real(g.Yshunt[s_]) * (VV + 2 * V[s_, n - 1].real)
                    if s_ in phase_map:
                        acc2 = 0
                        for x in range(n):
                            acc2 += np.conj(V[s_, x]) * phase_pp(s_, V[:, n - 1 - x])
                        cc_s -= acc2.real
                elif n == 1:
                    cc_s = Pi[s_] - np.real(g.Yshunt[s_])
                    if s_ in phase_map:
                        for v_ in phase_map[s_][1]:
                            cc_s -= v_.real
                else:
                    PP = _spmv_rows(g, np.array([s_]), V[:, 1])[0]
                    slack_CC[1] = PP
                    cc_s = 0 - (np.conj(V[s_, 1]) * PP).real
                    if g.conduc[s_]:
                        cc_s -= np.real(g.Yshunt[s_]) * 2 * V[s_, 1].real
                    if s_ in phase_map:
                        acc2 = 0
                        for x in range(n):
                            acc2 += np.conj(V[s_, x]) * phase_pp(s_, V[:, n - 1 - x])
                        cc_s -= acc2.real
                rhs[2 * N] = np.real(cc_s)
            rhs_hist[:, n] = rhs
            if pv_bus_model == 1:                                 # helm.py:584-597
                sub = np.zeros(length)
                for i in list_gen:
                    for r_, v_ in pv1_cols[int(i)].items():
                        sub[r_] += Vre_PV[i, n] * v_
                x_ = solve(rhs - sub)
            else:
                x_ = solve(rhs)                                   # helm.py:602
            coeff[:, n] = x_
            V[:, n] = x_[0:2 * N:2] + 1j * x_[1:2 * N:2]          # helm.py:412-420
            if pv_bus_model == 1:
                V[pv, n] = Vre_PV[pv, n] + 1j * x_[2 * pv + 1]
            aux = np.zeros(N, dtype=complex)                      # helm.py:423-433
            for k in range(n):
                aux = aux + W[:, k] * V[:, n - k]
            W[:, n] = -aux
            # ---- convergence schedule (helm.py:608-641)
            flag_mismatch = False
            m += 1
            if (m - 1) % 2 == 0 and m > 3:
                p1 = pade_batch(V, m - 2)
                p2 = pade_batch(V, m)
                mag1, rad1 = np.abs(p1), np.angle(p1)
                mag2, rad2 = np.abs(p2), np.angle(p2)
                bad = (np.abs(mag1 - mag2) > mismatch) | (np.abs(rad1 - rad2) > mismatch)
                flag_mismatch = bool(bad.any())
                if not flag_mismatch:
                    Vprof[:] = p2
                    if enforce_Q_limits:
                        # check_PVLIM_violation / Q_iny (helm.py:468-490, 452-465)
                        viol = False
                        Vre, Vim = Vprof.real, Vprof.imag
                        newly = []
                        for i in list_gen:
                            q = 0
                            for p in range(g.ptr[i], g.ptr[i + 1]):
                                k = g.idx[p]
                                yre, yim = g.yfull[p].real, g.yfull[p].imag
                                q += Vim[i] * (yre * Vre[k] - yim * Vim[k]) \
                                    - Vre[i] * (yre * Vim[k] + yim * Vre[k])
                            qg = q + Qd[i]
                            Qg[i] = qg
                            if qg > g.Qgmax[i] or qg < g.Qgmin[i]:
                                viol = True
                                btype[i] = PQ
                                newly.append(int(i))
                                Qg[i] = g.Qgmax[i] if qg > g.Qgmax[i] else g.Qgmin[i]
                        if viol:
                            switched += newly
                            list_gen = np.setdiff1d(list_gen, newly, assume_unique=True)
                            list_coef.append(m)
                            recalc = True
                    if not recalc:
                        list_coef.append(m)
                    if record:
                        passes.append(dict(A=A, V=V[:, :m].copy(), W=W[:, :m].copy(),
                                           coeff=coeff[:, :m].copy(), rhs=rhs_hist[:, :m].copy(),
                                           btype=None, m=m))
                    break
            if m > max_coef - 1:                                  # helm.py:654
                diverged = True
                break
        if not recalc or diverged:
            break
    info = {'list_coef': list_coef, 'switched': switched, 'passes': passes}
    if diverged:
        return None, info
    if dsb:
        info['Ploss'] = pade(coeff[2 * N], m)                     # helm.py:963-965
    return Vprof.copy(), info


def polar(V):
    """(|V|, angle rad) as cmath.polar does per element (helm.py:615-617)."""
    return np.abs(V), np.angle(V)
