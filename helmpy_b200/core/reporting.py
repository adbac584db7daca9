"""Result layout of ``helm``: branch flows, power balance, printers and the
``Results <algorithm> <name> <scale> <mismatch>.xlsx/.txt`` writers
(reference helmpy/core/helm.py:662-850).  Host code: it runs once per solve on
O(N_branches) data and only when ``detailed_run_print`` or ``save_results`` is set
(helm.py:962).  The .xlsx is written with the stdlib writer in ``xlsx.py`` (the
reference's ``pd.ExcelWriter.save`` no longer exists in pandas >= 2).

Reference quirk kept on purpose (file parity of the report for scale != 1):
``helm`` resets the case scale *before* the report (helm.py:957-958), so the
"demanded power" lines use the unscaled loads while generation uses the scaled
``run.Pg``.
"""

import numpy as np

from . import xlsx

BRANCH_HEADERS = ['From Bus', 'To Bus', 'From-To P injection (MW)', 'From-To Q injection (MVAR)',
                  'To-From P injection (MW)', 'To-From Q injection (MVAR)',
                  'P flow through branch and elements (MW)', 'Q flow through branch and elements (MVAR)']


def convert_complex_to_polar_voltages(complex_voltage, N=None):
    """|V| and angle in degrees (helm.py:662-667)."""
    v = np.asarray(complex_voltage)
    out = np.empty((len(v), 2), dtype=np.float64)
    out[:, 0] = np.absolute(v)
    out[:, 1] = np.angle(v, deg=True)
    return out


def branch_flows(case, V):
    """Power_branches [N_branches, 8] (helm.py:685-712), x100 MVA base."""
    nb = case.N_branches
    f = np.array([b[0] for b in case.Ybr_list], dtype=np.int64)
    t = np.array([b[1] for b in case.Ybr_list], dtype=np.int64)
    Y = np.array([b[2] for b in case.Ybr_list], dtype=np.complex128).reshape(nb, 2, 2)
    Vf, Vt = V[f], V[t]
    If = Y[:, 0, 0] * Vf + Y[:, 0, 1] * Vt
    It = Y[:, 1, 0] * Vf + Y[:, 1, 1] * Vt
    S_ft = Vf * np.conj(If) * 100
    S_tf = Vt * np.conj(It) * 100
    S_el = S_ft + S_tf
    P = np.zeros((nb, 8), dtype=np.float64)
    P[:, 0], P[:, 1] = f, t
    P[:, 2], P[:, 3] = S_ft.real, S_ft.imag
    P[:, 4], P[:, 5] = S_tf.real, S_tf.imag
    P[:, 6], P[:, 7] = S_el.real, S_el.imag
    return P


def _injection(case, V, i):
    """Complex power injected at bus i: V_i conj(sum_k Y_ik V_k) (helm.py:436-465)."""
    p0, p1 = case.adj_ptr[i], case.adj_ptr[i + 1]
    I = np.sum(case.Y_val[p0:p1] * V[case.adj_idx[p0:p1]])
    return V[i] * np.conj(I)


def power_balance(case, V, Qg, Pg_run, list_gen, enforce_Q_limits, algorithm, K=None):
    """(Power_branches, S_gen, S_load, S_mismatch, Pmismatch) as helm.py:670-759."""
    P = branch_flows(case, V)
    P_losses_line = np.sum(P[:, 6]) / 100
    Q_losses_line = np.sum(P[:, 7]) * 1j / 100
    nz = case.Shunt != 0
    S_shunt = np.sum(V[nz] * np.conj(V[nz] * case.Shunt[nz])) if nz.any() else 0
    Qload = np.sum(case.Qd) * 1j
    Pload = np.sum(case.Pd)
    Qg = np.array(Qg, dtype=np.float64)
    slack = case.slack
    if not enforce_Q_limits:
        for i in list_gen:
            Qg[i] = _injection(case, V, i).imag + case.Qd[i]          # Q_iny(i) + Qd[i], helm.py:730
    s_slack = _injection(case, V, slack)
    Qgen = (np.sum(Qg) + s_slack.imag + case.Qd[slack]) * 1j
    Pmismatch = None
    if 'DS' in algorithm:
        Pmismatch = P_losses_line + np.real(S_shunt)
        Pgen = np.sum(Pg_run + K * Pmismatch)
    else:
        Pgen = np.sum(Pg_run) + s_slack.real + case.Pd[slack]
    S_gen = (Pgen + Qgen) * 100
    S_load = (Pload + Qload) * 100
    S_mismatch = (P_losses_line + Q_losses_line + S_shunt) * 100
    return P, S_gen, S_load, S_mismatch, Pmismatch


def print_voltage_profile(V_polar_final, N):
    """helm.py:762-775."""
    print("\n\tVoltage profile:")
    print("   Bus    Magnitude (p.u.)    Phase Angle (degrees)")
    fmt = "{:>6d}\t     {:1.6f}\t\t{:11.6f}"
    if N <= 31:
        print(*(fmt.format(i, mag, ang) for i, (mag, ang) in enumerate(V_polar_final)), sep='\n')
    else:
        print(*(fmt.format(i, mag, ang) for i, (mag, ang) in enumerate(V_polar_final[0:14])), sep='\n')
        print(*3 * ("     .\t         .\t\t      .",), sep='\n')
        print(*(fmt.format(i, mag, ang) for i, (mag, ang) in enumerate(V_polar_final[N - 14:N], N - 14)), sep='\n')
    print()


def create_power_balance_string(mismatch, scale, algorithm, list_coef_or_iterations, S_gen, S_load,
                                S_mismatch, Ploss=None, Pmismatch=None):
    """helm.py:778-807 (same text, same number formats)."""
    what = 'Coefficients' if algorithm[0:2] == 'HE' else 'Iterations'
    out = ('Scale: {}   Mismatch: {}'.format(scale, mismatch)
           + '   {:s} per PVLIM-PQ switches: {:s}'.format(what, str(list_coef_or_iterations))
           + "\n\n  *  Power Balance:  *"
           + "\n\nTotal generated power (MVA):  ----------------> {:< 22.15f} {:=+23.15f} j".format(np.real(S_gen), np.imag(S_gen))
           + "\nTotal demanded power (MVA):  -----------------> {:< 22.15f} {:=+23.15f} j".format(np.real(S_load), np.imag(S_load))
           + "\nTotal power through branches and shunt"
           + "\nelements (mismatch) (MVA):  ------------------> {:< 22.15f} {:=+23.15f} j".format(np.real(S_mismatch), np.imag(S_mismatch))
           + "\n\nComparison: Generated power (MVA):  ----------> {:< 22.15f} {:=+23.15f} j".format(np.real(S_gen), np.imag(S_gen))
           + "\n            Demanded plus mismatch power (MVA): {:< 22.15f} {:=+23.15f} j".format(
               np.real(S_load + S_mismatch), np.imag(S_load + S_mismatch)))
    if Ploss is not None:
        out += ("\n\nComparison: Active power losses 'Ploss' variable (MW):  ---------------------> {:< 22.15f}".format(np.real(Ploss * 100))
                + "\n            Active power through branches and shunt elements 'Pmismatch' (MW): {:< 22.15f}".format(np.real(Pmismatch * 100)))
    return out


def write_results_on_files(mismatch, scale, algorithm, V, V_polar_final, Power_branches,
                           results_file_name, power_balance_string):
    """helm.py:810-850: '<files_name>.xlsx' (sheets Buses, Branches) and '<files_name>.txt'."""
    files_name = 'Results' + ' ' + algorithm + ' ' + str(results_file_name) + ' ' + str(scale) + ' ' + str(mismatch)
    n = len(V)
    buses = ('Buses', ['Complex Voltages', 'Voltages Magnitude', 'Voltages Phase Angle'], list(range(n)),
             [[str(complex(v)) for v in V], list(V_polar_final[:, 0]), list(V_polar_final[:, 1])])
    branches = ('Branches', BRANCH_HEADERS, list(range(Power_branches.shape[0])),
                [list(Power_branches[:, k]) for k in range(8)])
    xlsx_name = files_name + '.xlsx'
    xlsx.write_xlsx(xlsx_name, [buses, branches])
    txt_name = files_name + ".txt"
    with open(txt_name, "w") as f:
        f.write(power_balance_string)
    print("\nResults have been written on the files:\n\t%s \n\t%s" % (xlsx_name, txt_name))
    return xlsx_name, txt_name


def pade(series, m):
    """[L/L] Pade approximant at s = 1 of a scalar series (analytic_continuation.py:29-50); used
    here only for the Ploss line of the distributed-slack report (helm.py:963-965)."""
    s = np.asarray(series[:m], dtype=complex)
    L = (len(s) - 1) // 2
    C = np.array([[s[r + c + 1] for c in range(L)] for r in range(L)], dtype=complex)
    b = np.ones(L + 1, dtype=complex)
    b[1:] = (-np.linalg.solve(C, s[L + 1:len(s)]))[::-1]
    a = np.array([sum(s[k] * b[i - k] for k in range(i + 1)) for i in range(L + 1)])
    return np.sum(a) / np.sum(b)


def k_factors(case, Pg_run, list_gen, DSB_model):
    """compute_k_factor / K_slack_1 (helm.py:493-525)."""
    K = np.zeros(case.N)
    if not DSB_model:
        K[case.slack] = 1
        return K
    contrib = [i for i in list_gen if Pg_run[i] > 0]
    if Pg_run[case.slack] > 0:
        contrib.append(case.slack)
    tot = sum(Pg_run[i] for i in contrib)
    for i in contrib:
        K[i] = Pg_run[i] / tot
    return K


def report(case, V, Qg, bus_types, scale, mismatch, list_coef, algorithm, enforce_Q_limits,
           detailed_run_print, save_results, results_file_name, DSB_model=False, ploss_series=None):
    """What helm() does after convergence when printing or saving (helm.py:962-984)."""
    Pg_run = case.Pg * scale                       # run.Pg: copy taken while the case was scaled
    list_gen = [int(i) for i in np.nonzero(np.asarray(bus_types) == 1)[0]]
    K = Ploss = None
    if 'DS' in algorithm:
        if DSB_model:                              # helm.py:506 with classes.py:286
            Pg_run[case.slack] = np.sum(case.Pd * scale) - np.sum(case.Pg * scale)
        K = k_factors(case, Pg_run, list_gen, DSB_model)
        Ploss = pade(ploss_series, list_coef[-1])
    Pb, S_gen, S_load, S_mis, Pmis = power_balance(case, V, Qg, Pg_run, list_gen, enforce_Q_limits, algorithm, K)
    polar = convert_complex_to_polar_voltages(V)
    text = create_power_balance_string(mismatch, scale, algorithm, list_coef, S_gen, S_load, S_mis, Ploss, Pmis)
    if detailed_run_print:
        print_voltage_profile(polar, case.N)
        print(text)
    if save_results:
        return write_results_on_files(mismatch, scale, algorithm, V, polar, Pb,
                                      results_file_name if results_file_name is not None else case.name, text)
    return None
