"""HELM solver entry points — host side of the drop-in boundary.

``helm(case, ...)`` keeps the reference's signature, return convention and
unconditional prints (reference helmpy/core/helm.py:896-985, 651, 655);
``helm_batch`` is the new scenario-batch call (B independent ``helm`` runs of
one grid at different load scales).  All arithmetic of the run —
matrix assembly, numeric LU, per-order right-hand sides, triangular solves,
V/W recurrences, Pade continuation, convergence test, Q-limit switching and
the restart loop — executes in the CUDA library behind the C ABI
(include/helm_b200.h).  There is no CPU path: without the built library or
without a CUDA device these functions raise ``HelmB200Error``.

All ten variants of the reference's test matrix run on the GPU: pv_bus_model 1 / 2,
classic slack, distributed slack methods 1 / 2 with or without participation
factors (test/test_helmpy.py:111-122).  The default (PV2, classic slack) uses the
tuned kernels; the others use straightforward variant kernels
(csrc/helm_variants.cuh).

Documented deviation: the reference's argument validation
prints its message and then runs anyway because ``(False, None)`` is truthy
(helm.py:881-911); here the message is printed and ``None`` is returned.
"""

import numpy as np

from . import _lib
from ._lib import HELM_MAX_PASSES, ST_CONVERGED


def validate_arguments(case, detailed_run_print, mismatch, scale, max_coefficients,
                       enforce_Q_limits, results_file_name, save_results, pv_bus_model,
                       DSB_model, DSB_model_method):
    """Same checks and messages as reference helm.py:853-892; returns a bool."""
    if (type(detailed_run_print) is not bool or type(mismatch) is not float
            or not (type(scale) is float or type(scale) is int)
            or type(max_coefficients) is not int or type(enforce_Q_limits) is not bool
            or not (results_file_name is None or type(results_file_name) is str)
            or type(save_results) is not bool or type(pv_bus_model) is not int
            or type(DSB_model) is not bool
            or not (DSB_model_method is None or type(DSB_model_method) is int)):
        print("Erroneous argument type.")
        return False
    if max_coefficients < 5:
        print("'max_coefficients' must be equal or greater than five (5).")
        return False
    if pv_bus_model not in (1, 2):
        print("'pv_bus_model' must be the integer 1 or 2.",)
        return False
    if DSB_model_method is not None and DSB_model_method not in (1, 2):
        print("'DSB_model_method' must be the integer 1 or 2.",)
        return False
    return True


def _case_fingerprint(case):
    """Cheap digest of everything of ``case`` the plan uploads (tens of KB)."""
    import hashlib
    h = hashlib.blake2b(digest_size=16)
    for name in ("Pd", "Qd", "Pg", "Qgmax", "Qgmin", "V", "Yshunt", "Ytrans_val", "Y_val", "phase_val",
                 "adj_ptr", "adj_idx", "phase_ptr", "phase_idx"):
        h.update(np.ascontiguousarray(getattr(case, name)).tobytes())
    h.update(np.ascontiguousarray(case.bus_types_code()).tobytes())
    return h.digest()


def get_plan(case):
    """Symbolic analysis + device-resident grid for ``case``; cached on the object.

    The reference rebuilds its run state from ``case`` on every call (helm.py:934), so
    whatever the caller did to the case in between (``set_scale``, edited loads or
    limits) is what it solves.  The cached plan is therefore keyed on a digest of the
    arrays it uploaded and rebuilt when they changed.
    """
    fp = _case_fingerprint(case)
    plan = getattr(case, "_b200_plan", None)
    if plan is None or getattr(case, "_b200_plan_fp", None) != fp:
        plan = _lib.Plan(case)
        case._b200_plan = plan
        case._b200_plan_fp = fp
    return plan


class BatchResult:
    """Result of ``helm_batch``: voltages [B, N] complex128 (rows of scenarios
    that did not converge are zero), status [B], coefficients per pass
    (``run.list_coef`` of each scenario), PV->PQ switch counts, solver stats."""

    def __init__(self, V, status, ncoef, nswitch, stats):
        self.V = V
        self.status = status
        self.ncoef = ncoef
        self.nswitch = nswitch
        self.stats = stats

    @property
    def converged(self):
        return self.status == ST_CONVERGED

    def list_coef(self, b):
        row = self.ncoef[b]
        return [int(x) for x in row[row > 0]]


def helm_batch(case, scales, mismatch=1e-4, max_coefficients=100, enforce_Q_limits=True,
               group=0, use_graph=True, poll_every=0, max_resident=0,
               pv_bus_model=2, DSB_model=False, DSB_model_method=None, outages=None,
               record_switches=False):
    """Solve ``len(scales)`` load-scaling scenarios of ``case`` on the current GPU.

    Scenario b equals ``helm(case, scale=scales[b], mismatch=..., ...)``
    (set_scale semantics of reference classes.py:215-219).  ``case`` is not
    mutated.  Host buffers in, host buffers out.  At most ``max_resident`` scenarios
    (default 4096) hold GPU state at a time; the rest of the batch queues and takes
    over slots as scenarios finish, so device memory does not grow with the batch.

    N-1 scenarios: ``outages[b]`` = row of the Branches sheet removed in scenario b
    (-1: none); equals ``helm`` on a case loaded without that branch.  Outages that
    island part of the grid (``non_islanding_branches`` lists the safe ones) end with
    status 3/4, as a singular system would.
    """
    plan = get_plan(case)
    if DSB_model and DSB_model_method is None:
        DSB_model_method = 2                                   # helm.py:913-914
    prm = _lib.Plan.make_params(mismatch, max_coefficients, enforce_Q_limits, group, use_graph,
                                poll_every, max_resident, pv_bus_model, DSB_model_method or 0,
                                DSB_model, record_switches)
    if scales is None:
        scales = np.ones(len(outages))
    V, status, ncoef, nswitch, stats = plan.solve_host(np.asarray(scales, dtype=np.float64), prm, outages)
    return BatchResult(V, status, ncoef, nswitch, stats)


def helm(case, detailed_run_print=False, mismatch=1e-4, scale=1, max_coefficients=100,
         enforce_Q_limits=True, results_file_name=None, save_results=False, pv_bus_model=2,
         DSB_model=False, DSB_model_method=None):
    """Drop-in for reference ``helmpy.helm`` (helm.py:896-985).

    Returns the complex bus voltages (ndarray complex128, Buses-sheet order) or
    ``None`` when the maximum number of coefficients is reached.
    """
    if not validate_arguments(case, detailed_run_print, mismatch, scale, max_coefficients,
                              enforce_Q_limits, results_file_name, save_results, pv_bus_model,
                              DSB_model, DSB_model_method):
        return None
    if DSB_model and DSB_model_method is None:
        DSB_model_method = 2
    pv_str = 'PV1' if pv_bus_model == 1 else 'PV2'             # helm.py:916-922
    if DSB_model_method is not None:
        algorithm = 'HELM DS ' + ('M1' if DSB_model_method == 1 else 'M2') + ' ' + pv_str
    else:
        algorithm = 'HELM ' + pv_str
    res = helm_batch(case, [float(scale)], mismatch=mismatch, max_coefficients=max_coefficients,
                     enforce_Q_limits=enforce_Q_limits, group=1, pv_bus_model=pv_bus_model,
                     DSB_model=DSB_model, DSB_model_method=DSB_model_method,
                     record_switches=detailed_run_print)
    coefs = res.list_coef(0)
    converged = res.status[0] == ST_CONVERGED
    if detailed_run_print:
        # the run's progress lines, replayed in the reference's order (helm.py:573, 489, 647)
        switches = get_plan(case).last_switches(1, res.nswitch)[0]
        npass = len(coefs)
        for p, m in enumerate(coefs):
            for n in range(1, m):
                print("Computing coefficient: %d" % n)
            for (sp, bus, q) in switches:
                if sp == p:
                    lim = case.Qgmax[bus] if q > case.Qgmax[bus] else case.Qgmin[bus]
                    print('Bus %d exceeded its Qgen limit with %f. The exceeded limit %f will be assigned to the bus'
                          % (bus + 1, q, lim))
            if p < npass - 1 or not converged:
                print("At coefficient %d the system is to be resolved due to PVLIM to PQ switches\n" % m)
        if res.status[0] == _lib.ST_DIVERGED:
            for n in range(1, max_coefficients):
                print("Computing coefficient: %d" % n)
    if not converged:
        print('\nMaximum number of coefficients has been reached. The problem has no physical solution')
        return None
    print('\nConvergence has been reached. %d coefficients were calculated' % coefs[-1])
    V = res.V[0].copy()
    if detailed_run_print or save_results:
        from . import reporting
        plan = get_plan(case)
        Qg, bus_types = plan.last_state(1)
        ploss = None
        if DSB_model_method is not None:
            ploss = plan.last_ploss(1, max_coefficients, DSB_model_method, pv_bus_model)[0]
        reporting.report(case, V, Qg[0], bus_types[0], scale=scale, mismatch=mismatch, list_coef=coefs,
                         algorithm=algorithm, enforce_Q_limits=enforce_Q_limits,
                         detailed_run_print=detailed_run_print, save_results=save_results,
                         results_file_name=results_file_name, DSB_model=DSB_model, ploss_series=ploss)
    return V


def non_islanding_branches(case):
    """Rows of the Branches sheet whose removal keeps the grid connected (no bridge).

    A branch with a parallel twin is never a bridge; the rest are checked with one
    DFS low-link pass over the bus graph (host code, O(N + branches)).
    """
    import sys
    f = np.asarray(case.branch_contrib['f'])
    t = np.asarray(case.branch_contrib['t'])
    N = case.N
    mult = {}
    for b in range(len(f)):
        key = (min(f[b], t[b]), max(f[b], t[b]))
        mult[key] = mult.get(key, 0) + 1
    adj = [[] for _ in range(N)]
    for (a, c) in mult:
        if a != c:
            adj[a].append(c)
            adj[c].append(a)
    disc = [-1] * N
    low = [0] * N
    bridges = set()
    timer = 0
    for root in range(N):
        if disc[root] != -1:
            continue
        stack = [(root, -1, 0)]
        disc[root] = low[root] = timer
        timer += 1
        while stack:
            v, parent, idx = stack.pop()
            if idx < len(adj[v]):
                stack.append((v, parent, idx + 1))
                u = adj[v][idx]
                if u == parent:
                    continue
                if disc[u] == -1:
                    disc[u] = low[u] = timer
                    timer += 1
                    stack.append((u, v, 0))
                else:
                    low[v] = min(low[v], disc[u])
            else:
                if parent >= 0:
                    low[parent] = min(low[parent], low[v])
                    if low[v] > disc[parent]:
                        bridges.add((min(v, parent), max(v, parent)))
    out = []
    for b in range(len(f)):
        key = (min(f[b], t[b]), max(f[b], t[b]))
        if mult[key] > 1 or key not in bridges:
            out.append(b)
    return np.asarray(out, dtype=np.int32)
