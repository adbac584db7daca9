"""Newton-Raphson cross-check entry points: ``nr`` and ``nr_ds``.

Same signatures, prints, return convention (complex voltages or ``None`` on
divergence) and iteration semantics as the reference (helmpy/core/nr.py:506-598,
nr_ds.py:547-646).  NR is *only* the cross-check of the HELM path (north star:
"NR remains the cross-check"; SURVEY §2 rows 9-10 put it out of scope as a
kernel), so it runs on the host — but with a sparse Jacobian (scipy) instead of
the reference's dense N x N one (nr.py:284), so it scales to the 10k-bus
configurations.

Algorithm as in the reference: polar form, unknowns d(theta) for every
non-reference bus and dV/V for every PQ bus (nr.py:244-282, update
``V *= 1 + dV``, nr.py:390-406); flat start with generator set-points
(nr.py:191-230); converged when max|dP, dQ| <= Mismatch (nr.py:343-386);
after convergence every PVLIM bus whose Qg = Qi + Qd violates a limit becomes
PQ at the limit and the iteration continues from the current voltages
(nr.py:409-436).  ``nr_ds`` adds the unknown Ploss shared through
participation factors K (nr_ds.py:246-291, 353-377, 395, 434).
"""

import numpy as np
from scipy.sparse import csr_matrix, diags, hstack, vstack
from scipy.sparse.linalg import spsolve

from .classes import create_case_data_object_from_xlsx
from . import reporting

PQ, PV, REF = 0, 1, 2

last_run = {}   # list_iterations, Ploss, K of the most recent call (the reference keeps them in globals)


def _newton(case, Mismatch, Scale, MaxIterations, Enforce_Qlimits, distributed, DSB_model, detailed):
    N = case.N
    rows = np.repeat(np.arange(N), np.diff(case.adj_ptr))
    Y = csr_matrix((case.Y_val, (rows, case.adj_idx)), shape=(N, N))
    Pd = case.Pd * Scale
    Qd = case.Qd * Scale
    Pesp = case.Pg * Scale
    btype = case.bus_types_code().astype(np.int8)
    slack = int(case.slack)
    Vm = np.ones(N)
    gen = btype != PQ
    Vm[gen] = case.V[gen]
    th = np.zeros(N)
    Qg = np.zeros(N)
    Ploss = 0.0
    list_iterations = []
    iterations = 0
    K = np.zeros(N)
    while True:
        Pg = Pesp.copy()
        if distributed:                                     # nr_ds.py:353-377
            Pg[slack] = np.sum(Pd) - np.sum(Pg)
            K = np.zeros(N)
            if DSB_model:
                contrib = np.nonzero((btype != PQ) & (Pg > 0))[0]
                K[contrib] = Pg[contrib] / np.sum(Pg[contrib])
            else:
                K[slack] = 1.0
        pvpq = np.nonzero(btype != REF)[0]
        pq = np.nonzero(btype == PQ)[0]
        prow = np.arange(N) if distributed else pvpq        # dP equations
        diverged = False
        while True:
            V = Vm * np.exp(1j * th)
            I = Y @ V
            S = V * np.conj(I)
            dP = Pg + K * Ploss - Pd - S.real               # nr_ds.py:395 / nr.py:359
            dQ = Qg - Qd - S.imag
            F = np.concatenate([dP[prow], dQ[pq]])
            err = np.max(np.abs(F)) if len(F) else 0.0
            if err <= Mismatch:
                if detailed:
                    print("Program converged with a maximum error of:", err)
                break
            if detailed:
                print("Maximum error:", err)
            iterations += 1
            if iterations == MaxIterations + 1:             # nr.py:378-381
                print("\nIteration number: %d. It is assumed that the program diverged." % (iterations - 1))
                diverged = True
                break
            if detailed:
                print("\nIteration number: %d" % iterations)
            Vd = diags(V)
            dS_dth = 1j * Vd @ (diags(np.conj(I)) - (Y @ Vd).conj())
            dS_dv = Vd @ (Y @ Vd).conj() + diags(np.conj(I) * V)      # d S / d|V| * |V|
            dS_dth, dS_dv = csr_matrix(dS_dth), csr_matrix(dS_dv)
            J11 = dS_dth.real[prow][:, pvpq]
            J12 = dS_dv.real[prow][:, pq]
            J21 = dS_dth.imag[pq][:, pvpq]
            J22 = dS_dv.imag[pq][:, pq]
            if distributed:
                kcol = csr_matrix(np.concatenate([K[prow], np.zeros(len(pq))])[:, None])
                J = hstack([kcol, vstack([hstack([J11, J12]), hstack([J21, J22])])]).tocsc()
            else:
                J = vstack([hstack([J11, J12]), hstack([J21, J22])]).tocsc()
            dx = spsolve(J, F)
            if distributed:
                Ploss -= dx[0]                              # nr_ds.py:434
                dx = dx[1:]
            th[pvpq] += dx[:len(pvpq)]
            Vm[pq] = Vm[pq] * (1 + dx[len(pvpq):])          # nr.py:403
        if diverged:
            last_run.update(list_iterations=list_iterations, Ploss=Ploss, K=K, Qg=Qg, btype=btype, Pg=Pg)
            return None
        V = Vm * np.exp(1j * th)
        if not Enforce_Qlimits:
            print("Convergence has been reached")
            list_iterations.append(iterations)
            break
        list_iterations.append(iterations)                  # nr.py:413-414
        iterations = 0
        S = V * np.conj(Y @ V)
        restart = False
        for i in np.nonzero(btype == PV)[0]:                # nr.py:420-434
            Qg[i] = S[i].imag + Qd[i]
            if Qg[i] > case.Qgmax[i] or Qg[i] < case.Qgmin[i]:
                before = Qg[i]
                restart = True
                btype[i] = PQ
                Qg[i] = case.Qgmax[i] if Qg[i] > case.Qgmax[i] else case.Qgmin[i]
                if detailed:
                    print('PVLIM bus %d exceeded its reactive power generation limit at %f MVAR. '
                          'Exceeded limit: %f MVAR' % (i, before * 100, Qg[i] * 100))
        if not restart:
            print("Convergence has been reached")
            break
    last_run.update(list_iterations=list_iterations, Ploss=Ploss, K=K, Qg=Qg, btype=btype, Pg=Pg)
    return V


def _check_types(Print_Details, Mismatch, Scale, MaxIterations, Enforce_Qlimits, Results_FileName):
    ok = (type(Print_Details) is bool and type(Mismatch) is float and type(Results_FileName) is str
          and (type(Scale) is float or type(Scale) is int) and type(MaxIterations) is int
          and type(Enforce_Qlimits) is bool)
    if not ok:
        print("Erroneous argument type.")       # nr.py:519-531
    return ok


def _report(case, V, algorithm, Scale, Mismatch, Enforce_Qlimits, Print_Details, Save_results, name):
    P = reporting.branch_flows(case, V)
    polar = reporting.convert_complex_to_polar_voltages(V)
    if Print_Details:
        reporting.print_voltage_profile(polar, case.N)
    if Save_results:
        text = 'Scale: {}   Mismatch: {}   Iterations per PVLIM-PQ switches: {}'.format(
            Scale, Mismatch, last_run['list_iterations'])
        reporting.write_results_on_files(Mismatch, Scale, algorithm, V, polar, P, name, text)


def nr(grid_data_file_path, Print_Details=False, Mismatch=1e-4, Scale=1, MaxIterations=15,
       Enforce_Qlimits=True, Results_FileName='', Save_results=False):
    """Newton-Raphson power flow (reference nr.py:506-598)."""
    if not _check_types(Print_Details, Mismatch, Scale, MaxIterations, Enforce_Qlimits, Results_FileName):
        return None
    case = grid_data_file_path if hasattr(grid_data_file_path, 'adj_ptr') \
        else create_case_data_object_from_xlsx(grid_data_file_path)
    V = _newton(case, Mismatch, Scale, MaxIterations, Enforce_Qlimits, False, False, Print_Details)
    if V is not None and (Print_Details or Save_results):
        _report(case, V, 'NR', Scale, Mismatch, Enforce_Qlimits, Print_Details, Save_results,
                Results_FileName or case.name)
    return V


def nr_ds(grid_data_file_path, Print_Details=False, Mismatch=1e-4, Scale=1, MaxIterations=15,
          Enforce_Qlimits=True, DSB_model=True, Results_FileName='', Save_results=False):
    """Newton-Raphson with the distributed slack bus model (reference nr_ds.py:547-646)."""
    if not _check_types(Print_Details, Mismatch, Scale, MaxIterations, Enforce_Qlimits, Results_FileName) \
            or type(DSB_model) is not bool:
        if type(DSB_model) is not bool:
            print("Erroneous argument type.")
        return None
    case = grid_data_file_path if hasattr(grid_data_file_path, 'adj_ptr') \
        else create_case_data_object_from_xlsx(grid_data_file_path)
    V = _newton(case, Mismatch, Scale, MaxIterations, Enforce_Qlimits, True, DSB_model, Print_Details)
    if V is not None and (Print_Details or Save_results):
        _report(case, V, 'NR DS', Scale, Mismatch, Enforce_Qlimits, Print_Details, Save_results,
                Results_FileName or case.name)
    return V


class NRBatchResult:
    """Result of ``nr_batch``: ``V`` [B, N] complex128 (zeros where not converged), ``status``
    (2 converged, 3 diverged, 4 singular Jacobian), ``iterations`` [B, 16] Newton steps per
    Q-limit pass (the reference's ``list_iterations``), ``nswitch`` PV->PQ switches."""

    def __init__(self, V, status, iterations, nswitch):
        self.V, self.status, self.iterations, self.nswitch = V, status, iterations, nswitch

    @property
    def converged(self):
        return self.status == 2

    def list_iterations(self, b):
        row = self.iterations[b]
        return [int(x) - 1 for x in row[row > 0]]            # the C ABI stores steps + 1 (0 = pass not run)


def nr_batch(case, scales, Mismatch=1e-4, MaxIterations=15, Enforce_Qlimits=True):
    """Newton-Raphson cross-check for ``len(scales)`` load-scaling scenarios on the GPU.

    Scenario b equals ``nr(case, Mismatch=..., Scale=scales[b], ...)`` (reference nr.py:506-598,
    classic slack): same flat start, same update, same convergence test and Q-limit handling; the
    Jacobian lives on the 2x2-block pattern of the HELM plan and is factorized and solved by the
    same kernels (csrc/helm_nr.cuh).  ``case`` is not mutated.
    """
    from .helm import get_plan
    plan = get_plan(case)
    V, status, iters, nswitch = plan.nr_host(np.asarray(scales, dtype=np.float64), Mismatch, MaxIterations,
                                             Enforce_Qlimits)
    return NRBatchResult(V, status, iters, nswitch)
