"""Reader / writer for MATPOWER case files (``function mpc = caseN`` ... ``mpc.bus = [...]``).

The reference's grids are MATPOWER cases dumped to .xlsx by a MATLAB helper
(reference helmm/Obtain_case_MATPOWER.m:1-16: ``mpc.bus``, ``mpc.branch``, ``mpc.gen``
written to the sheets Buses / Branches / Generators without a header).  This module
reads the ``.m`` text directly, so cases that were never converted (case2383wp,
case9241pegase, case13659pegase, PGLib files) can be loaded without MATLAB, and writes
the reference's .xlsx layout so the unmodified reference can read the same grid.

The reference hard-codes a 100 MVA base (classes.py:137-139); a file with another
``mpc.baseMVA`` is rescaled to that base on load (powers and shunts are in MW / MVAr,
branch impedances in p.u. on the file's base).
"""

import re
from os.path import basename, splitext

import numpy as np

from .classes import BASE_MVA, create_case_data_object_from_arrays
from . import xlsx


def _matrix(text, field):
    m = re.search(r'mpc\.%s\s*=\s*\[(.*?)\]\s*;' % re.escape(field), text, flags=re.S)
    if not m:
        raise ValueError("MATPOWER case has no mpc.%s matrix" % field)
    rows = []
    for line in m.group(1).replace(';', '\n').split('\n'):
        line = line.split('%', 1)[0].strip()
        if line:
            rows.append([float(x) for x in re.split(r'[\s,]+', line) if x])
    width = max(len(r) for r in rows)
    return np.array([r + [0.0] * (width - len(r)) for r in rows], dtype=np.float64)


def read_matpower_tables(path):
    """Return (buses, branches, generators, baseMVA) of a MATPOWER ``.m`` case file."""
    with open(path) as f:
        text = f.read()
    text = re.sub(r'%\{.*?%\}', '', text, flags=re.S)          # block comments
    mb = re.search(r'mpc\.baseMVA\s*=\s*([0-9.eE+\-]+)\s*;', text)
    base = float(mb.group(1)) if mb else BASE_MVA
    return _matrix(text, 'bus'), _matrix(text, 'branch'), _matrix(text, 'gen'), base


def create_case_data_object_from_matpower(path, case_name=None):
    """CaseData from a MATPOWER ``.m`` file (same columns as the reference's .xlsx sheets,
    reference classes.py:126-158; generators that are switched off are kept, as the reference does)."""
    buses, branches, gens, base = read_matpower_tables(path)
    if base != BASE_MVA:
        k = BASE_MVA / base                     # per-unit values move to the 100 MVA base
        branches = branches.copy()
        branches[:, 2:4] *= k                   # r, x
        branches[:, 4] /= k                     # line charging b
    if branches.shape[1] < 10:                  # no tap / shift columns: plain lines
        branches = np.hstack([branches, np.zeros((branches.shape[0], 10 - branches.shape[1]))])
    if case_name is None:
        case_name = splitext(basename(path))[0]
    return create_case_data_object_from_arrays(buses, branches, gens, case_name)


def matpower_to_xlsx(m_path, xlsx_path):
    """Write the reference's .xlsx layout (sheets Buses / Branches / Generators, no header) for a
    MATPOWER ``.m`` case: the stdlib equivalent of reference helmm/Obtain_case_MATPOWER.m."""
    buses, branches, gens, _ = read_matpower_tables(m_path)
    xlsx.write_numeric_xlsx(xlsx_path, [('Buses', buses), ('Branches', branches), ('Generators', gens)])
