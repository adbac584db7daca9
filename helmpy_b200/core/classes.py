"""Grid model and .xlsx case loader — host side of the drop-in boundary.

Mirrors the *interface* of the reference's ``CaseData`` and
``create_case_data_object_from_xlsx`` (reference helmpy/core/classes.py:116-225)
but stores the network sparsely: the reference keeps two dense N x N
complex128 matrices (classes.py:203-204), which is 3 GB each at 13 659 buses.
Here the series-admittance Laplacian ``Ytrans``, the full ``Y`` and the
phase-shifter corrections live as CSR arrays on the (sorted) bus adjacency
``branches_buses``; ``case.Ytrans`` / ``case.Y`` / ``case.Yre`` / ``case.Yimag``
remain available as lazily-built dense views for code written against the
reference.

Arithmetic follows the reference branch by branch (classes.py:9-113) so that
every admittance is bit-identical to the reference's: same per-unit base 100
(classes.py:137-139), same accumulation order for parallel branches, the same
two formulas for "tap in {0,1}" and "tap otherwise" (classes.py:24, 64-108).

Deliberate deviation: for a phase shifter whose *to* bus
already has a phase entry for a different neighbour, the reference appends the
new entry to the *from* bus's lists (classes.py:47-49, 89-91).  No shipped
case triggers it; this loader files the entry under the *to* bus, as intended.
"""

from os.path import basename

import numpy as np

from . import xlsx

BASE_MVA = 100.0  # hard-coded in the reference (classes.py:137-139, 156-158)


class CaseData:
    """Data of a case (attribute-compatible with reference classes.py:180-225)."""

    def __init__(self, name, N, N_generators):
        self.name = name
        self.N = N
        self.N_branches = 0
        self.slack_bus = 0
        self.slack = 0
        self.Number_bus = dict()
        self.Buses_type = ['PQ' for _ in range(N)]
        self.list_gen = np.empty(max(N_generators - 1, 0), dtype=int)
        self.V = np.ones(N, dtype=np.float64)
        self.Pd = np.zeros(N, dtype=np.float64)
        self.Qd = np.zeros(N, dtype=np.float64)
        self.Pg = np.zeros(N, dtype=np.float64)
        self.Qgmax = np.zeros(N, dtype=np.float64)
        self.Qgmin = np.zeros(N, dtype=np.float64)
        self.Shunt = np.zeros(N, dtype=np.complex128)
        self.conduc_buses = np.full(N, False)
        self.Yshunt = np.zeros(N, dtype=np.complex128)
        self.branches_buses = [[i] for i in range(N)]
        self.Ybr_list = list()
        self.phase_barras = np.full(N, False)
        self.phase_dict = dict()
        self.scale = 1
        # sparse storage (CSR over branches_buses); filled by _finalize()
        self.adj_ptr = None      # int32 [N+1]
        self.adj_idx = None      # int32 [nnz], sorted per row, includes the diagonal
        self.Ytrans_val = None   # complex128 [nnz]
        self.Y_val = None        # complex128 [nnz]  (Ytrans + diag(Yshunt) + phase)
        self.phase_ptr = None    # int32 [N+1]
        self.phase_idx = None    # int32 [nph]
        self.phase_val = None    # complex128 [nph]
        self.branch_table = None  # float64 [N_branches, >=10] raw Branches sheet
        # per-branch contributions (for N-1 outage scenarios): what each branch added to Ytrans
        # (series admittance y), to Yshunt at its two ends and to the phase-shifter terms
        self.branch_contrib = None  # dict of arrays: f, t, y, sh_f, sh_t, ph_ft, ph_tf
        self._dense = {}

    # ---- reference API (classes.py:215-225) ---------------------------------
    def set_scale(self, scale):
        self.scale = scale
        self.Pd *= scale
        self.Qd *= scale
        self.Pg *= scale

    def reset_scale(self):
        self.Pd /= self.scale
        self.Qd /= self.scale
        self.Pg /= self.scale
        self.scale = 1

    # ---- dense compatibility views -------------------------------------------
    def _densify(self, key, vals):
        if key not in self._dense:
            M = np.zeros((self.N, self.N), dtype=np.complex128)
            rows = np.repeat(np.arange(self.N), np.diff(self.adj_ptr))
            M[rows, self.adj_idx] = vals
            self._dense[key] = M
        return self._dense[key]

    @property
    def Ytrans(self):
        return self._densify('Ytrans', self.Ytrans_val)

    @property
    def Y(self):
        return self._densify('Y', self.Y_val)

    @property
    def Yre(self):
        return np.real(self.Y)

    @property
    def Yimag(self):
        return np.imag(self.Y)

    # ---- helpers --------------------------------------------------------------
    def bus_types_code(self):
        """uint8 per bus: 0 = PQ, 1 = PV/PVLIM, 2 = Slack."""
        code = np.zeros(self.N, dtype=np.uint8)
        for i, t in enumerate(self.Buses_type):
            if t == 'Slack':
                code[i] = 2
            elif t != 'PQ':
                code[i] = 1
        return code


def _stamp_branches(case, branches, skip_branch=None):
    """Accumulate branch admittances (reference classes.py:9-113).

    Sparse accumulators keyed by (row, col) replace the dense matrices; the
    order of the floating-point additions is the reference's.
    """
    N = case.N
    ytr = {}
    phase = {}      # bus -> ([nbrs], [values]) in insertion order
    adj = case.branches_buses
    Yshunt = case.Yshunt

    def acc(d, i, j, v):
        d[(i, j)] = d.get((i, j), 0j) + v

    def add_phase(bus, nbr, v):
        if bus in phase:
            lst = phase[bus]
            if nbr in lst[0]:
                lst[1][lst[0].index(nbr)] += v
            else:
                lst[0].append(nbr)
                lst[1].append(v)
        else:
            phase[bus] = [[nbr], [v]]
            case.phase_barras[bus] = True

    nb = branches.shape[0]
    contrib = dict(f=np.zeros(nb, dtype=np.int32), t=np.zeros(nb, dtype=np.int32),
                   y=np.zeros(nb, dtype=np.complex128), sh_f=np.zeros(nb, dtype=np.complex128),
                   sh_t=np.zeros(nb, dtype=np.complex128), ph_ft=np.zeros(nb, dtype=np.complex128),
                   ph_tf=np.zeros(nb, dtype=np.complex128))
    case.branch_contrib = contrib
    for b in range(nb):
        f = case.Number_bus[int(branches[b, 0])]
        t = case.Number_bus[int(branches[b, 1])]
        ybr = np.zeros((2, 2), dtype=np.complex128)
        case.Ybr_list.append([f, t, ybr])
        contrib['f'][b], contrib['t'][b] = f, t
        if skip_branch is not None and b == skip_branch:
            continue   # branch outage: contributes nothing, adjacency not created
        z = branches[b, 2] + 1j * branches[b, 3]
        half_b = 1j * branches[b, 4] / 2
        tap = branches[b, 8]
        shift_deg = branches[b, 9]
        untapped = (tap == 0 or tap == 1)
        tap_inv = 1.0 if untapped else 1 / tap
        if z != 0:
            ys = 1 / z
            y = ys if untapped else ys * tap_inv
            if shift_deg == 0:
                ybr[0, 1] = ybr[1, 0] = -y
            else:
                sh = np.deg2rad(shift_deg)
                y_ft = y / (np.exp(-1j * sh))
                y_tf = y / (np.exp(1j * sh))
                ybr[0, 1] = -y_ft
                ybr[1, 0] = -y_tf
                add_phase(f, t, y - y_ft)
                add_phase(t, f, y - y_tf)
                contrib['ph_ft'][b], contrib['ph_tf'][b] = y - y_ft, y - y_tf
            acc(ytr, f, t, -y)
            acc(ytr, f, f, y)
            acc(ytr, t, f, -y)
            acc(ytr, t, t, y)
            contrib['y'][b] = y
        else:
            ys = y = 0
        if untapped:
            ybr[0, 0] = ybr[1, 1] = half_b + y
            Yshunt[f] += half_b
            Yshunt[t] += half_b
            contrib['sh_f'][b] = contrib['sh_t'][b] = half_b
        else:
            sh_f = (ys + half_b) * (tap_inv * tap_inv)
            sh_t = ys + half_b
            ybr[0, 0] = sh_f
            ybr[1, 1] = sh_t
            Yshunt[f] += sh_f - y
            Yshunt[t] += sh_t - y
            contrib['sh_f'][b], contrib['sh_t'][b] = sh_f - y, sh_t - y
        if t not in adj[f]:
            adj[f].append(t)
        if f not in adj[t]:
            adj[t].append(f)

    for i in range(N):
        adj[i].sort()
    case.phase_dict = phase
    return ytr


def _finalize(case, ytr):
    N = case.N
    adj = case.branches_buses
    ptr = np.zeros(N + 1, dtype=np.int32)
    for i in range(N):
        ptr[i + 1] = ptr[i] + len(adj[i])
    idx = np.empty(ptr[N], dtype=np.int32)
    yt = np.zeros(ptr[N], dtype=np.complex128)
    for i in range(N):
        row = adj[i]
        idx[ptr[i]:ptr[i + 1]] = row
        for p, j in enumerate(row):
            yt[ptr[i] + p] = ytr.get((i, j), 0j)
    yfull = yt.copy()
    pptr = np.zeros(N + 1, dtype=np.int32)
    pidx, pval = [], []
    for i in range(N):
        if case.Yshunt[i].real != 0:
            case.conduc_buses[i] = True
        row = adj[i]
        yfull[ptr[i] + row.index(i)] += case.Yshunt[i]
        if case.phase_barras[i]:
            nbrs, vals = case.phase_dict[i]
            for k, v in zip(nbrs, vals):
                yfull[ptr[i] + row.index(k)] += v
                pidx.append(k)
                pval.append(v)
        pptr[i + 1] = len(pidx)
    case.adj_ptr, case.adj_idx = ptr, idx
    case.Ytrans_val, case.Y_val = yt, yfull
    case.phase_ptr = pptr
    case.phase_idx = np.asarray(pidx, dtype=np.int32)
    case.phase_val = np.asarray(pval, dtype=np.complex128)
    case._dense = {}


def create_case_data_object_from_arrays(buses, branches, generators, case_name='case',
                                        skip_branch=None):
    """Build a CaseData from the three MATPOWER-layout tables (no header).

    Column use follows reference classes.py:126-158 and 12-18.
    """
    buses = np.asarray(buses, dtype=np.float64)
    branches = np.asarray(branches, dtype=np.float64)
    generators = np.asarray(generators, dtype=np.float64)
    N = buses.shape[0]
    n_gen = generators.shape[0]
    case = CaseData(case_name, N, n_gen)
    case.N_branches = branches.shape[0]
    case.branch_table = branches
    case.Pd[:] = buses[:, 2] / BASE_MVA
    case.Qd[:] = buses[:, 3] / BASE_MVA
    case.Shunt[:] = buses[:, 5] * 1j / BASE_MVA + buses[:, 4] / BASE_MVA
    case.Yshunt[:] = case.Shunt
    # the reference leaves V / Qgmax / Qgmin of non-generator buses uninitialised
    # (np.empty, classes.py:194-199); they are never read for PQ buses.
    for i in range(N):
        num = int(buses[i, 0])
        case.Number_bus[num] = i
        if buses[i, 1] == 3:
            case.slack_bus = num
            case.slack = i
    pos = 0
    list_gen = np.empty(n_gen, dtype=int)
    for g in range(n_gen):
        bus_i = case.Number_bus[int(generators[g, 0])]
        if bus_i != case.slack:
            list_gen[pos] = bus_i
            pos += 1
        case.Buses_type[bus_i] = 'PVLIM'
        case.V[bus_i] = generators[g, 5]
        case.Pg[bus_i] = generators[g, 1] / BASE_MVA
        case.Qgmax[bus_i] = generators[g, 3] / BASE_MVA
        case.Qgmin[bus_i] = generators[g, 4] / BASE_MVA
    case.list_gen = list_gen[:pos].copy()
    case.Buses_type[case.slack] = 'Slack'
    case.Pg[case.slack] = 0
    ytr = _stamp_branches(case, branches, skip_branch=skip_branch)
    _finalize(case, ytr)
    return case


def create_case_data_object_from_xlsx(grid_data_file_path, case_name=None):
    """Creates a case data object from an .xlsx (reference classes.py:116-177).

    Sheets ``Buses``, ``Branches``, ``Generators``; no header row; MATPOWER
    column order.  Error convention of the reference: message + ``None``.
    """
    if not (case_name is None or type(case_name) is str):
        print("Erroneous argument type.")
        return None
    if case_name is None:
        case_name = basename(grid_data_file_path[0:-5])
    buses = xlsx.read_numeric_sheet(grid_data_file_path, 'Buses')
    branches = xlsx.read_numeric_sheet(grid_data_file_path, 'Branches')
    generators = xlsx.read_numeric_sheet(grid_data_file_path, 'Generators')
    return create_case_data_object_from_arrays(buses, branches, generators, case_name)
