"""CPU: pin the oracle (oracle/helm_oracle.py) against the reference's golden
vectors and against committed outputs of the unmodified reference."""
import numpy as np
import pytest

from common import CASES, TOL_ANG, TOL_MAG, golden_results, load_case, polar_err, ref_outputs
from oracle import helm_oracle as ho

FAST = ['case9', 'case118', 'case1354pegase']


@pytest.mark.parametrize('name', CASES)
def test_oracle_matches_reference_golden_xlsx(name):
    """Reference test/test_helmpy.py:73-92: helm(mismatch=1e-8, scale=1) vs
    'Results HELM <case> 1 1e-08.xlsx' (|V| and angle in degrees)."""
    case = load_case(name)
    V, info = ho.helm_oracle(case, mismatch=1e-8, scale=1, max_coefficients=40)
    g = golden_results(name)
    assert V is not None
    assert np.max(np.abs(np.abs(V) - g['classic_mag'])) < 1e-11
    assert np.max(np.abs(np.angle(V, deg=True) - g['classic_ang_deg'])) < 1e-9


@pytest.mark.parametrize('name', CASES)
def test_oracle_matches_reference_runs(name):
    """Same inputs as the committed reference runs: same pass structure, same
    switched buses, voltages to round-off; divergence reproduced as None."""
    case = load_case(name)
    runs, _ = ref_outputs(name)
    if name == 'case2869pegase':
        runs = [r for r in runs if r['scale'] in (1.0, 1.2)]
    for r in runs:
        V, info = ho.helm_oracle(case, mismatch=1e-8, scale=r['scale'], max_coefficients=40)
        assert (V is not None) == r['converged']
        assert info['list_coef'] == r['list_coef']
        assert info['switched'] == r['switched']
        if V is not None:
            dmag, dang = polar_err(V, r['V'])
            assert dmag < 1e-11 and dang < 1e-11


@pytest.mark.parametrize('name', ['case9', 'case118'])
def test_oracle_series_coefficients(name):
    """Per-order parity: coefficient columns of the last pass (run.coefficients)."""
    case = load_case(name)
    _, z = ref_outputs(name)
    V, info = ho.helm_oracle(case, mismatch=1e-8, scale=1, max_coefficients=40, record=True)
    ref = z['last_pass_coefficients']
    mine = info['passes'][-1]['coeff']
    assert mine.shape == ref.shape
    assert np.max(np.abs(mine - ref)) < 1e-10
    assert np.max(np.abs(info['passes'][-1]['W'] - z['last_pass_W'])) < 1e-9


VARIANTS = [(2, False, None), (1, False, None), (2, False, 1), (2, False, 2), (1, False, 1), (1, False, 2),
            (2, True, 1), (2, True, 2), (1, True, 1), (1, True, 2)]       # test/test_helmpy.py:111-122


@pytest.mark.parametrize('name', FAST)
@pytest.mark.parametrize('variant', VARIANTS)
def test_oracle_variants_match_reference_golden_xlsx(name, variant):
    """The reference's full test matrix (test/test_helmpy.py:103-158): classic-slack runs against
    'Results HELM <case> 1 1e-08.xlsx', distributed-slack runs (scale 1.02) against
    'Results HELM DS <case> 1.02 1e-08.xlsx'."""
    pv, dsb, method = variant
    case = load_case(name)
    V, info = ho.helm_oracle(case, mismatch=1e-8, scale=1.02 if dsb else 1, max_coefficients=40,
                             pv_bus_model=pv, DSB_model=dsb, DSB_model_method=method)
    g = golden_results(name)
    key = 'ds' if dsb else 'classic'
    assert V is not None
    assert np.max(np.abs(np.abs(V) - g[key + '_mag'])) < 1e-11
    assert np.max(np.abs(np.angle(V, deg=True) - g[key + '_ang_deg'])) < 1e-9


@pytest.mark.reference
@pytest.mark.parametrize('variant', VARIANTS[1:])
def test_oracle_variants_against_live_reference(variant):
    from oracle import ref_shim
    pv, dsb, method = variant
    rc = ref_shim.ref_case('/root/reference/data/cases/case118.xlsx')
    case = load_case('case118')
    for s in (1.0, 1.1):
        Vr, ir = ref_shim.ref_helm(rc, scale=s, pv_bus_model=pv, DSB_model=dsb, DSB_model_method=method)
        Vo, io = ho.helm_oracle(case, mismatch=1e-8, scale=s, max_coefficients=40, pv_bus_model=pv,
                                DSB_model=dsb, DSB_model_method=method)
        assert ir['list_coef'] == io['list_coef'] and ir['switched'] == io['switched']
        assert np.max(np.abs(Vr - Vo)) < 1e-11


def test_pade_equals_epsilon_and_known_answer():
    """Pade [L/L](1) of a geometric series is exact: sum r^k = 1/(1-r)."""
    r = 0.5 + 0.3j
    s = r ** np.arange(9)
    assert abs(ho.pade(s, 9) - 1 / (1 - r)) < 1e-12
    assert abs(ho.epsilon(s, 3) - 1 / (1 - r)) < 1e-12
    rng = np.random.default_rng(0)
    s = (rng.standard_normal(15) + 1j * rng.standard_normal(15)) * 0.6 ** np.arange(15)
    assert abs(ho.pade(s, 15) - ho.epsilon(s, 15)) < 1e-9
    assert abs(ho.pade_batch(s[None, :], 15)[0] - ho.pade(s, 15)) < 1e-13


def test_golden_vs_matpower():
    """Third oracle (reference test/Compare with MATPOWER.py:28-37)."""
    for name in CASES:
        g = golden_results(name)
        case = load_case(name)
        ang = g['matpower_ang_deg'] - g['matpower_ang_deg'][case.slack]
        assert np.max(np.abs(g['classic_mag'] - g['matpower_mag'])) < 1e-8
        assert np.max(np.abs(g['classic_ang_deg'] - ang)) < 1e-6


@pytest.mark.reference
def test_oracle_against_live_reference():
    from oracle import ref_shim
    path = '/root/reference/data/cases/case118.xlsx'
    rc = ref_shim.ref_case(path)
    case = load_case('case118')
    for s in (0.9, 1.1):
        Vr, ir = ref_shim.ref_helm(rc, scale=s)
        Vo, io = ho.helm_oracle(case, mismatch=1e-8, scale=s, max_coefficients=40)
        assert ir['list_coef'] == io['list_coef'] and ir['switched'] == io['switched']
        dmag, dang = polar_err(Vo, Vr)
        assert dmag < 1e-11 and dang < 1e-11
