"""The reference's package name: ``import helmpy`` / ``from helmpy.core import helm`` in an
unmodified user script resolve to the B200 build (reference helmpy/__init__.py:1,
helmpy/core/__init__.py:1-4).  Run in a fresh interpreter so that the reference package, which other
tests import under the same name, cannot interfere."""
import os
import subprocess
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(code, timeout=600):
    env = dict(os.environ, PYTHONPATH=REPO + os.pathsep + os.path.join(REPO, 'tests'))
    out = subprocess.run([sys.executable, '-c', code], env=env, capture_output=True, text=True, timeout=timeout)
    assert out.returncode == 0, out.stderr[-3000:]
    return out.stdout


def test_helmpy_names_resolve_to_the_b200_build():
    out = _run(
        "import sys, helmpy, helmpy_b200\n"
        "from helmpy.core import helm, nr, nr_ds, create_case_data_object_from_xlsx\n"
        "import helmpy.core\n"
        "assert helmpy.IMPLEMENTATION == 'helmpy_b200'\n"
        "for n in ('helm', 'nr', 'nr_ds', 'create_case_data_object_from_xlsx'):\n"
        "    assert getattr(helmpy, n) is getattr(helmpy_b200, n), n\n"
        "    assert getattr(helmpy.core, n) is getattr(helmpy_b200, n), n\n"
        "assert callable(helmpy.core.helm) and not isinstance(helmpy.core.helm, type(sys))\n"   # the function, as in the reference
        "assert sys.modules['helmpy.core.helm'].helm is helm\n"
        "from helmpy.core.helm import helm as h2\n"
        "from helmpy.core.nr_ds import nr_ds as d2\n"
        "from helmpy.core.classes import CaseData\n"
        "assert h2 is helm and d2 is nr_ds and helmpy.core.helm is helm\n"
        "print('ok')\n")
    assert out.strip().endswith('ok')


@pytest.mark.gpu
def test_unmodified_user_script_runs_on_the_gpu():
    """The reference's own test flow (test/test_helmpy.py:24, 75): load a case, call helmpy.helm."""
    out = _run(
        "import numpy as np, helmpy\n"
        "from common import load_case, golden_results\n"
        "case = load_case('case118')\n"
        "V = helmpy.helm(case, mismatch=1e-8, max_coefficients=40)\n"
        "g = golden_results('case118')\n"
        "assert np.max(np.abs(np.abs(V) - g['classic_mag'])) <= 1e-8\n"
        "from helmpy_b200.core import _lib\n"
        "assert _lib._lib is not None\n"
        "print('ok')\n")
    assert out.strip().endswith('ok')
