"""Generate the committed fixtures under tests/golden/ from the UNMODIFIED
reference (run in the build container, where /root/reference exists).

  tests/golden/cases/<case>.npz      Buses / Branches / Generators tables of
                                     the reference's data/cases/<case>.xlsx
  tests/golden/results_<case>.npz    |V| and angle[deg] columns of the reference's
                                     golden files data/results/Results HELM ... .xlsx
                                     (classic slack, scale 1; DS, scale 1.02) and
                                     of Results MATPOWER <case>.xlsx
  tests/golden/ref_<case>.npz        outputs of the reference helm() run here:
                                     V, list_coef, switched buses for several
                                     load scales (max_coefficients=40, quirk Q2),
                                     and the series coefficients of every pass
                                     at scale 1 for the two small cases

Usage:  python oracle/gen_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, REPO)

from oracle import ref_shim  # noqa: E402
from helmpy_b200.core import xlsx  # noqa: E402

CASES = ['case9', 'case118', 'case1354pegase', 'case2869pegase']
BRANCH_HEADERS = ['From Bus', 'To Bus', 'From-To P injection (MW)', 'From-To Q injection (MVAR)',
                  'To-From P injection (MW)', 'To-From Q injection (MVAR)',
                  'P flow through branch and elements (MW)', 'Q flow through branch and elements (MVAR)']
SCALES = {
    'case9': [1, 0.8, 0.9, 1.05, 1.2, 1.5],
    'case118': [1, 0.8, 0.95, 1.05, 1.2, 1.5],
    'case1354pegase': [1, 0.85, 1.03],
    'case2869pegase': [1, 0.85, 1.04, 1.2],
}


def main():
    out = os.path.join(REPO, 'tests', 'golden')
    os.makedirs(os.path.join(out, 'cases'), exist_ok=True)
    ref = ref_shim.REFERENCE_ROOT
    for name in CASES:
        path = os.path.join(ref, 'data', 'cases', name + '.xlsx')
        tables = {s.lower(): xlsx.read_numeric_sheet(path, s) for s in ('Buses', 'Branches', 'Generators')}
        np.savez_compressed(os.path.join(out, 'cases', name + '.npz'), **tables)

        res = {}
        for key, fname in (('classic', 'Results HELM %s 1 1e-08.xlsx' % name),
                           ('ds', 'Results HELM DS %s 1.02 1e-08.xlsx' % name)):
            cols = xlsx.read_table_with_header(os.path.join(ref, 'data', 'results', fname), 'Buses')
            res[key + '_mag'] = np.array(cols['Voltages Magnitude'], dtype=float)
            res[key + '_ang_deg'] = np.array(cols['Voltages Phase Angle'], dtype=float)
            br = xlsx.read_table_with_header(os.path.join(ref, 'data', 'results', fname), 'Branches')
            res[key + '_branches'] = np.array([br[h] for h in BRANCH_HEADERS], dtype=float).T
        mp = xlsx.read_numeric_sheet(os.path.join(ref, 'data', 'results', 'Results MATPOWER %s.xlsx' % name), 'Buses')
        res['matpower_mag'] = mp[:, 7]
        res['matpower_ang_deg'] = mp[:, 8]
        np.savez_compressed(os.path.join(out, 'results_%s.npz' % name), **res)

        if os.environ.get('GOLDEN_RESULTS_ONLY'):
            continue
        case = ref_shim.ref_case(path)
        d = {'scales': np.array(SCALES[name], dtype=float)}
        for k, s in enumerate(SCALES[name]):
            V, info = ref_shim.ref_helm(case, scale=s, record=(s == 1 and case.N < 200))
            d['conv_%d' % k] = np.array(V is not None)
            d['V_%d' % k] = V if V is not None else np.zeros(case.N, dtype=complex)
            d['list_coef_%d' % k] = np.array(info['list_coef'], dtype=int)
            d['switched_%d' % k] = np.array(info['switched'], dtype=int)
            if 'run' in info:
                run = info['run']
                m = info['list_coef'][-1]
                d['last_pass_coefficients'] = run.coefficients[:, :m].copy()
                d['last_pass_W'] = run.W[:, :m].copy()
            print(name, s, 'conv' if V is not None else 'None', info['list_coef'], len(info['switched']))
        np.savez_compressed(os.path.join(out, 'ref_%s.npz' % name), **d)
        if case.N < 200:      # the reference's own report text (helm.py:762-807), scale 1 and 1.05
            h = ref_shim.load_reference()
            import contextlib, io
            for sc in (1, 1.05):
                buf = io.StringIO()
                with contextlib.redirect_stdout(buf):
                    h.helm(case, detailed_run_print=True, mismatch=1e-8, scale=sc, max_coefficients=40)
                with open(os.path.join(out, 'ref_stdout_%s_%s.txt' % (name, sc)), 'w') as f:
                    f.write(buf.getvalue())
            for (pv, dm, me) in ((2, True, 2), (1, True, 1), (2, False, 1)):
                buf = io.StringIO()
                with contextlib.redirect_stdout(buf):
                    h.helm(case, detailed_run_print=True, mismatch=1e-8, scale=1.02, max_coefficients=40,
                           pv_bus_model=pv, DSB_model=dm, DSB_model_method=me)
                with open(os.path.join(out, 'ref_stdout_%s_ds_%d_%d_%d.txt' % (name, pv, int(dm), me)), 'w') as f:
                    f.write(buf.getvalue())


if __name__ == '__main__':
    main()
