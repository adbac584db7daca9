"""TEST INFRASTRUCTURE ONLY — runs the UNMODIFIED reference from /root/reference.

Used (a) to validate the numpy restatement in ``oracle/helm_oracle.py`` and
(b) by ``oracle/gen_golden.py`` to produce the fixtures under ``tests/golden/``.
/root/reference does not exist on the GPU box, so nothing that runs there may
import this module.  No product code may import anything under ``oracle/``.

Two shims are needed to run the reference in this image:
  1. ``numpy.int`` was removed in numpy 1.24 (reference classes.py:188-190);
  2. ``pandas.read_excel`` needs openpyxl, which is absent; it is replaced by
     the stdlib reader in ``helmpy_b200/core/xlsx.py`` wrapped in a DataFrame.
The reference sources themselves are imported in place, never copied.
"""

import contextlib
import io
import os
import sys

import numpy as np

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _REPO not in sys.path:
    sys.path.insert(0, _REPO)
# oracle/_ref: the unmodified reference package + grids as copied by oracle/build_ref.sh (git-ignored;
# it is what travels to the GPU box, where /root/reference does not exist)
LOCAL_COPY = os.path.join(_REPO, "oracle", "_ref")


def _default_root():
    env = os.environ.get("HELMPY_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isdir("/root/reference/helmpy/core"):
        return "/root/reference"
    return LOCAL_COPY


REFERENCE_ROOT = _default_root()


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "helmpy", "core"))


_ref = None


def load_reference():
    """Import the reference package (as module ``helmpy``) under the two shims."""
    global _ref
    if _ref is not None:
        return _ref
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    import pandas as pd
    from helmpy_b200.core import xlsx

    if not hasattr(np, "int"):
        np.int = int  # shim 1

    def read_excel(path, sheet_name=0, header=0, **_):  # shim 2
        if header is None:
            arr = xlsx.read_numeric_sheet(path, sheet_name)
            df = pd.DataFrame(arr)
            for c in df.columns:
                col = df[c].to_numpy()
                if np.all(col == np.floor(col)) and np.all(np.abs(col) < 2**53):
                    df[c] = col.astype(np.int64)
            return df
        cols = xlsx.read_table_with_header(path, sheet_name)
        return pd.DataFrame(cols)

    pd.read_excel = read_excel
    # the repository ships a `helmpy` package of its own (the drop-in name, bound to the CUDA
    # build): make sure the name resolves to the REFERENCE here, whatever was imported before
    for name in [m for m in sys.modules if m == "helmpy" or m.startswith("helmpy.")]:
        del sys.modules[name]
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        import helmpy  # noqa: the reference package
    finally:
        sys.path.remove(REFERENCE_ROOT)
    if getattr(helmpy, "IMPLEMENTATION", None) == "helmpy_b200":
        raise RuntimeError("imported the drop-in shim instead of the reference")
    _ref = helmpy
    return _ref


class Recorder:
    """Captures internals of one reference ``helm`` call."""

    def __init__(self):
        self.runs = []       # RunVariables instances created
        self.matrices = []   # CSC matrices handed to factorized()


@contextlib.contextmanager
def recording():
    """Rebind RunVariables / factorized in the reference's helm module (both
    are looked up as module globals at call time: helm.py:934, 106)."""
    load_reference()
    mod = sys.modules["helmpy.core.helm"]
    rec = Recorder()
    orig_rv, orig_fact = mod.RunVariables, mod.factorized

    class RV(orig_rv):
        def __init__(self, *a, **k):
            super().__init__(*a, **k)
            rec.runs.append(self)

    def fact(A):
        rec.matrices.append(A.copy())
        return orig_fact(A)

    mod.RunVariables, mod.factorized = RV, fact
    try:
        yield rec
    finally:
        mod.RunVariables, mod.factorized = orig_rv, orig_fact


def ref_case(path):
    h = load_reference()
    return h.create_case_data_object_from_xlsx(path)


def ref_helm(case, record=False, **kw):
    """Call the reference helm() quietly.  Always max_coefficients<=40 (quirk Q2).

    Returns (V or None, info dict with list_coef / switched buses / recorder).
    """
    h = load_reference()
    kw.setdefault("max_coefficients", 40)
    kw.setdefault("mismatch", 1e-8)
    buf = io.StringIO()
    with recording() as rec, contextlib.redirect_stdout(buf):
        try:
            V = h.helm(case, **kw)
        except Exception:
            if kw.get("scale", 1) != 1:
                case.reset_scale()
            raise
    run = rec.runs[-1]
    info = {
        "list_coef": list(run.list_coef),
        "switched": [int(i) for i in run.list_gen_remove],
        "stdout": buf.getvalue(),
    }
    if record:
        info["run"] = run
        info["matrices"] = rec.matrices
    return V, info


def ref_nr(path, **kw):
    h = load_reference()
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        return h.nr(path, **kw)
