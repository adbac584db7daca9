"""CPU: the Newton-Raphson cross-check (helmpy_b200.nr / nr_ds) against the
reference's golden files — the reference's own test (test/test_helmpy_nr.py:71-89)
compares NR with the HELM result files — and against the live reference."""
import contextlib
import io
import sys

import numpy as np
import pytest

from common import CASES, golden_results, load_case
from helmpy_b200.core.nr import last_run, nr, nr_ds

EXPECTED_ITERS = {'case9': [4], 'case118': [4, 3], 'case1354pegase': [5, 3, 2], 'case2869pegase': [5, 3, 3, 2]}


def _quiet(fn, *a, **k):
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        out = fn(*a, **k)
    return out, buf.getvalue()


@pytest.mark.parametrize('name', CASES)
def test_nr_matches_helm_golden(name):
    case = load_case(name)
    g = golden_results(name)
    V, out = _quiet(nr, case, Mismatch=1e-8)
    assert 'Convergence has been reached' in out
    assert last_run['list_iterations'] == EXPECTED_ITERS[name]      # SURVEY §6 probe of the reference
    assert np.max(np.abs(np.abs(V) - g['classic_mag'])) < 5e-9
    assert np.max(np.abs(np.angle(V, deg=True) - g['classic_ang_deg'])) < 1e-6
    # MATPOWER NR golden (angles re-referenced to the slack)
    ang = g['matpower_ang_deg'] - g['matpower_ang_deg'][case.slack]
    assert np.max(np.abs(np.abs(V) - g['matpower_mag'])) < 5e-9
    assert np.max(np.abs(np.angle(V, deg=True) - ang)) < 1e-6


@pytest.mark.parametrize('name', CASES)
def test_nr_ds_matches_helm_ds_golden(name):
    case = load_case(name)
    g = golden_results(name)
    V, _ = _quiet(nr_ds, case, Mismatch=1e-8, Scale=1.02, DSB_model=True)
    assert np.max(np.abs(np.abs(V) - g['ds_mag'])) < 5e-9
    assert np.max(np.abs(np.angle(V, deg=True) - g['ds_ang_deg'])) < 1e-6
    # classic slack through the distributed formulation equals plain NR
    V1, _ = _quiet(nr_ds, case, Mismatch=1e-8, DSB_model=False)
    V2, _ = _quiet(nr, case, Mismatch=1e-8)
    assert np.max(np.abs(V1 - V2)) < 1e-9


def test_nr_divergence_and_argument_errors():
    case = load_case('case2869pegase')
    V, out = _quiet(nr, case, Mismatch=1e-8, Scale=3.0, MaxIterations=6)
    assert V is None and 'It is assumed that the program diverged' in out
    V, out = _quiet(nr, case, Mismatch=1)
    assert V is None and 'Erroneous argument type.' in out


@pytest.mark.reference
@pytest.mark.parametrize('name', ['case9', 'case118'])
def test_nr_against_live_reference(name):
    from oracle import ref_shim
    h = ref_shim.load_reference()
    path = '/root/reference/data/cases/%s.xlsx' % name
    case = load_case(name)
    for fn_ref, fn, mod, kw in ((h.nr, nr, 'helmpy.core.nr', {}),
                                (h.nr_ds, nr_ds, 'helmpy.core.nr_ds', {'Scale': 1.02})):
        Vr, _ = _quiet(fn_ref, path, Mismatch=1e-8, **kw)
        Vr = np.array(Vr).copy()
        iters = list(sys.modules[mod].list_iterations)
        V, _ = _quiet(fn, case, Mismatch=1e-8, **kw)
        assert last_run['list_iterations'] == iters
        assert np.max(np.abs(V - Vr)) < 1e-12
