"""CPU: host logic — loader, xlsx round trip, symbolic analysis, C-ABI surface."""
import ctypes
import os
import re

import numpy as np
import pytest

import blocklu_emul as be
from common import CASES, load_case, load_tables
from helmpy_b200.core import _lib, xlsx
from helmpy_b200.core.classes import create_case_data_object_from_xlsx
from oracle import helm_oracle as ho

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    header = open(os.path.join(REPO, 'include', 'helm_b200.h')).read()
    declared = set(re.findall(r'\b(helm_[a-z_]+)\s*\(', header))
    assert declared == set(_lib.EXPORTED_SYMBOLS)
    for sym in declared:
        assert hasattr(lib, sym), sym


def test_xlsx_round_trip(tmp_path):
    b, br, g = load_tables('case118')
    p = str(tmp_path / 'case118.xlsx')
    xlsx.write_numeric_xlsx(p, [('Sheet1', np.zeros((0, 0))), ('Buses', b), ('Branches', br), ('Generators', g)])
    assert xlsx.sheet_names(p) == ['Sheet1', 'Buses', 'Branches', 'Generators']
    case = create_case_data_object_from_xlsx(p)
    ref = load_case('case118')
    assert case.name == 'case118' and case.N == 118
    assert np.array_equal(case.Ytrans_val, ref.Ytrans_val)
    assert np.array_equal(case.Y_val, ref.Y_val)
    assert case.branches_buses == ref.branches_buses
    assert create_case_data_object_from_xlsx(p, case_name=3) is None   # classes.py:118-120


def test_case_scale_semantics():
    case = load_case('case9')
    pd0 = case.Pd.copy()
    case.set_scale(1.1)
    assert np.allclose(case.Pd, pd0 * 1.1) and case.scale == 1.1
    case.reset_scale()
    assert np.allclose(case.Pd, pd0) and case.scale == 1


@pytest.mark.reference
@pytest.mark.parametrize('name', ['case9', 'case118', 'case1354pegase'])
def test_loader_bit_identical_to_reference(name):
    from oracle import ref_shim
    path = '/root/reference/data/cases/%s.xlsx' % name
    rc = ref_shim.ref_case(path)
    c = create_case_data_object_from_xlsx(path)
    assert np.array_equal(rc.Ytrans, c.Ytrans) and np.array_equal(rc.Y, c.Y)
    assert np.array_equal(rc.Yshunt, c.Yshunt)
    assert rc.branches_buses == c.branches_buses
    assert np.array_equal(rc.list_gen, c.list_gen) and rc.Buses_type == c.Buses_type
    for a in ('Pd', 'Qd', 'Pg'):
        assert np.array_equal(getattr(rc, a), getattr(c, a))
    assert set(rc.phase_dict) == set(c.phase_dict)
    for k in rc.phase_dict:
        assert rc.phase_dict[k][0] == c.phase_dict[k][0] and rc.phase_dict[k][1] == c.phase_dict[k][1]


@pytest.mark.parametrize('name', ['case9', 'case118', 'case1354pegase'])
def test_symbolic_structure(name):
    case = load_case(name)
    S = _lib.symbolic_arrays(case.adj_ptr, case.adj_idx)
    N = case.N
    assert sorted(S['perm']) == list(range(N))
    assert np.array_equal(S['iperm'][S['perm']], np.arange(N))
    # L rows strictly lower, ascending; rows of a level only depend on earlier levels
    lev = np.zeros(N, dtype=int)
    for l in range(len(S['llev_ptr']) - 1):
        lev[S['llev_ptr'][l]:S['llev_ptr'][l + 1]] = l
    for r in range(N):
        cols = S['lcol'][S['lptr'][r]:S['lptr'][r + 1]]
        assert np.all(np.diff(cols) > 0) and (len(cols) == 0 or cols[-1] < r)
        assert np.all(lev[cols] < lev[r])
    ulev = np.zeros(N, dtype=int)
    for l in range(len(S['ulev_ptr']) - 1):
        ulev[S['urow'][S['ulev_ptr'][l]:S['ulev_ptr'][l + 1]]] = l
    for q in range(N):
        r = S['urow'][q]
        cols = S['ucol'][S['uptr'][q]:S['uptr'][q + 1]]
        assert np.all(cols > r) and np.all(ulev[cols] < ulev[r])
    # structural symmetry of the fill: nL == nU
    assert len(S['lcol']) == len(S['ucol'])
    # every adjacency entry has its own slot
    assert len(set(S['adj_slot'])) == len(S['adj_slot'])


@pytest.mark.parametrize('name', ['case118', 'case1354pegase'])
def test_block_lu_static_pivoting_matches_superlu(name):
    """The fixed-pattern 2x2-block LU (no value pivoting) reproduces SuperLU's
    partial-pivoting solution on every Q-limit pass matrix of the case."""
    from scipy.sparse.linalg import splu
    case = load_case(name)
    S = _lib.symbolic_arrays(case.adj_ptr, case.adj_idx)
    g = ho.Grid(case)
    V, info = ho.helm_oracle(case, mismatch=1e-8, max_coefficients=40, record=True)
    bt = g.btype0.copy()
    done = 0
    for ps in info['passes']:
        A = ps['A']
        rhs = ps['rhs'][:, 3]
        xs = splu(A).solve(rhs)
        lu = be.factor(S, be.assemble(S, case.Ytrans_val, bt))
        x = be.solve(S, lu, rhs)
        assert np.max(np.abs(x - xs)) <= 1e-11 * np.max(np.abs(xs))
        # next pass: the buses switched so far become PQ
        nsw = len(info['switched'])
        done += 1
        # switched buses are recorded cumulatively; recover this pass's set from A of the next pass
        if done < len(info['passes']):
            An = info['passes'][done]['A'].tocsr()
            for i in range(case.N):
                if bt[i] == 1 and An.getrow(2 * i + 1).nnz > 1:
                    bt[i] = 0


@pytest.mark.parametrize('name', ['case118', 'case1354pegase'])
def test_factor_schedule_local_indices(name):
    """fdst_local addresses the same block as fdst inside the row buffer [L | D | U]."""
    case = load_case(name)
    S = _lib.symbolic_arrays(case.adj_ptr, case.adj_idx, case.bus_types_code())
    N = case.N
    nL = len(S['lcol'])
    for r in range(N):
        e0, e1 = S['lptr'][r], S['lptr'][r + 1]
        nl = e1 - e0
        base = S['dslot'][r]
        for e in range(e0, e1):
            for u in range(S['fptr'][e], S['fptr'][e + 1]):
                dl = S['fdst_local'][u]
                slot = e0 + dl if dl < nl else base + (dl - nl)
                assert slot == S['fdst'][u]
                assert S['fsrc'][u] >= nL          # sources are U blocks of an earlier row


@pytest.mark.parametrize('name', CASES)
def test_branch_flows_match_golden_results(name):
    """reporting.branch_flows on the golden voltages reproduces the 'Branches'
    sheet of the reference's result files (helm.py:685-712)."""
    from common import golden_results
    from helmpy_b200.core import reporting
    case = load_case(name)
    g = golden_results(name)
    V = g['classic_mag'] * np.exp(1j * np.deg2rad(g['classic_ang_deg']))
    P = reporting.branch_flows(case, V)
    ref = g['classic_branches']
    assert P.shape == ref.shape
    assert np.array_equal(P[:, :2], ref[:, :2])
    assert np.max(np.abs(P[:, 2:] - ref[:, 2:])) < 1e-6      # MW / MVAr


def test_non_islanding_branches():
    """Bridge detection for the N-1 sweep: removing a listed branch keeps the bus graph connected,
    removing an unlisted one splits it."""
    import helmpy_b200
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components
    case = load_case('case118')
    safe = set(int(b) for b in helmpy_b200.non_islanding_branches(case))
    f, t = case.branch_contrib['f'], case.branch_contrib['t']
    nb = len(f)
    for b in range(nb):
        keep = [k for k in range(nb) if k != b]
        A = coo_matrix((np.ones(len(keep)), (f[keep], t[keep])), shape=(case.N, case.N))
        ncomp, _ = connected_components(A, directed=False)
        assert (ncomp == 1) == (b in safe), b


@pytest.mark.parametrize('name', ['case9', 'case118', 'case1354pegase', 'case2869pegase', 'tiled'])
def test_flat_wave_schedule_solves(name):
    """The wave tables of k_solve_flat (symbolic.cpp 4b-flat), emulated lane by lane on the host,
    give the same x as the plain row-by-row block solve."""
    if name == 'tiled':         # several elimination subtrees side by side: long rows share levels
        import synth
        case = synth.tiled_case(['case1354pegase', 'case2869pegase', 'case118'])
    else:
        case = load_case(name)
    S = _lib.symbolic_arrays(case.adj_ptr, case.adj_idx, case.bus_types_code())
    N = case.N
    lptr, lcol, uptr, ucol, urow, dslot = S['lptr'], S['lcol'], S['uptr'], S['ucol'], S['urow'], S['dslot']
    nL = int(lptr[-1])
    nslots = nL + int(uptr[-1]) + N
    wt = S['wave_tab'].reshape(-1, 2)
    if (wt[-1, 1] & 63) == 0:       # padding entry (16-byte multiple for the bulk copy)
        wt = wt[:-1]
    meta, col = S['smeta'], S['slot_col']
    assert len(meta) == nslots and len(col) == nslots
    # tables of the shared-memory solve: {column, meta} per slot, chunks of consecutive waves
    scm = S['scm'].reshape(-1, 2)[:nslots]
    assert (scm[:, 0] == col).all() and (scm[:, 1] == meta).all()
    ch = S['solve_chunks'].reshape(-1, 4)
    assert ch[0, 0] == 0 and ch[0, 2] == 0 and (ch[:, 3] <= 128).all() and (ch[:, 3] >= 1).all()
    assert (ch[1:, 0] == ch[:-1, 0] + ch[:-1, 1]).all() and ch[-1, 0] + ch[-1, 1] == len(wt)
    assert (ch[1:, 2] == ch[:-1, 2] + ch[:-1, 3]).all() and ch[-1, 2] + ch[-1, 3] == nslots
    for w0, nwv, s0, ns in ch:
        assert wt[w0, 0] == s0 and int((wt[w0:w0 + nwv, 1] & 63).sum()) == ns
    rng = np.random.default_rng(3)
    lu = rng.standard_normal((nslots, 2, 2))
    rhs = rng.standard_normal((N, 2))
    # reference: rows in level order
    x = rhs.copy()
    for r in range(N):
        for e in range(lptr[r], lptr[r + 1]):
            x[r] -= lu[e] @ x[lcol[e]]
    for q in range(N):
        r = urow[q]
        base = nL + uptr[q] + q
        acc = x[r].copy()
        for t in range(uptr[q + 1] - uptr[q]):
            acc -= lu[base + 1 + t] @ x[ucol[uptr[q] + t]]
        x[r] = lu[base] @ acc
    # emulation of the waves
    covered = np.zeros(nslots, dtype=int)
    y = np.full((N, 2), np.nan)
    nol = 0
    while nol < N and lptr[nol + 1] == lptr[nol]:
        nol += 1
    y[:nol] = rhs[:nol]
    carry = None
    carry_d = None
    final_u = np.zeros(N, dtype=bool)
    this_level = set()          # rows written since the last SYNC wave: must not be gathered
    cont_pending = False
    for e0, info in wt:
        n = info & 63
        is_u = bool(info & 256)
        assert 1 <= n <= 32
        assert (info & 64) or not (info & 128), 'a wave that hands on a carry must be ordered (SYNC)'
        if info & 64:
            this_level = set()
        else:
            assert not cont_pending, 'the wave after a continued row must be ordered after it'
        cont_pending = False
        covered[e0:e0 + n] += 1
        lane = 0
        while lane < n:
            m = meta[e0 + lane]
            assert m & 32, 'a row piece must start with a head lane'
            span = m & 31
            assert ((info >> 9) & 31) >= span and lane + span < n
            r = m >> 9
            v = np.zeros(2)
            for k in range(lane, lane + span + 1):
                mk = meta[e0 + k]
                assert (mk & 31) == lane + span - k and (mk >> 9) == r
                if not (mk & 256):
                    assert not np.isnan(y[col[e0 + k]]).any(), 'gather of a value that is not yet final'
                    assert col[e0 + k] not in this_level, 'gather inside the level that writes the value'
                    assert final_u[col[e0 + k]] or not is_u
                    v += lu[e0 + k] @ y[col[e0 + k]]
            if m & 64:
                b = y[r] if is_u else rhs[r]
                d = lu[e0 + lane] if is_u else None
                if is_u:
                    assert m & 256
            else:
                b, d = carry, carry_d
            acc = b - v
            if m & 128:
                y[r] = d @ acc if is_u else acc
                this_level.add(r)
                final_u[r] = is_u
            else:
                assert info & 128 and lane + span == n - 1
                carry, carry_d = acc, d
                cont_pending = True
            lane += span + 1
    assert (covered == 1).all()
    assert np.allclose(y, x, rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize('name', ['case9', 'case118', 'case1354pegase', 'case2869pegase', 'tiled', 'long_rows'])
def test_row_lane_schedule_solves(name):
    """The packed per-lane records of k_solve_rows (symbolic.cpp 4b-rows), emulated lane by lane on the
    host: every L entry, pivot and U entry is used exactly once, no lane gathers a value written inside
    its own level, and x equals the plain row-by-row block solve."""
    if name == 'tiled':
        import synth
        case = synth.tiled_case(['case1354pegase', 'case2869pegase', 'case118'])
    elif name == 'long_rows':       # 13.6 k buses: rows of up to 120 blocks, cut into pieces (BASELINE config 5's grid)
        import synth
        case = synth.tiled_case(['case2869pegase'] * 4 + ['case1354pegase'] + ['case118'] * 7)
    else:
        case = load_case(name)
    S = _lib.symbolic_arrays(case.adj_ptr, case.adj_idx, case.bus_types_code())
    N = case.N
    lptr, lcol, uptr, ucol, urow = S['lptr'], S['lcol'], S['uptr'], S['ucol'], S['urow']
    nL = int(lptr[-1])
    nslots = nL + int(uptr[-1]) + N
    if name == 'long_rows':
        assert int(np.diff(lptr).max()) > 96
    desc = S['rl_desc'].reshape(-1, 32, 4)
    assert len(desc) >= 2
    nw = len(desc) - 2                      # the last two records are the idle look-ahead waves
    assert (desc[nw:, :, 0] == -1).all() and (desc[nw:, :, 3] == (0xFFFF | (1 << 30))).all()     # ... and say stop
    assert (desc[nw:, :, 2] == 0).all() and not (desc[:nw, :, 3] & (1 << 30)).any()
    rng = np.random.default_rng(5)
    lu = rng.standard_normal((nslots, 2, 2))
    rhs = rng.standard_normal((N, 2))
    x = rhs.copy()
    for r in range(N):
        for e in range(lptr[r], lptr[r + 1]):
            x[r] -= lu[e] @ x[lcol[e]]
    for q in range(N):
        r = urow[q]
        base = nL + uptr[q] + q
        acc = x[r].copy()
        for t in range(uptr[q + 1] - uptr[q]):
            acc -= lu[base + 1 + t] @ x[ucol[uptr[q] + t]]
        x[r] = lu[base] @ acc
    covered = np.zeros(nslots, dtype=int)
    y = np.full((N, 2), np.nan)
    nol = 0
    while nol < N and lptr[nol + 1] == lptr[nol]:
        nol += 1
    y[:nol] = rhs[:nol]
    qof = S['qof']
    left_l = np.diff(lptr).astype(int)                      # blocks of the row still to be summed, forward
    left_u = np.array([uptr[qof[r] + 1] - uptr[qof[r]] + 1 for r in range(N)], dtype=int)   # backward, with the pivot
    started = np.zeros(N, dtype=bool)                       # a piece of the row has run in the current phase
    this_level = set()
    slot_col = S['slot_col']
    seen_u = False
    npieces = 0
    for w in range(nw):
        d = desc[w]
        flags = d[:, 3]
        is_u = bool(flags[0] & (1 << 26))
        assert ((flags >> 25) & 3 == (flags[0] >> 25) & 3).all() and ((flags >> 27) & 7 == (flags[0] >> 27) & 7).all()
        assert is_u or not seen_u
        if is_u and not seen_u:
            assert (left_l == 0).all()
            started[:] = False
        seen_u = is_u
        left = left_u if is_u else left_l
        rounds = (int(flags[0]) >> 27) & 7
        if flags[0] & (1 << 25):
            this_level = set()
        part = np.zeros((32, 2))
        nblk = np.zeros(32, dtype=int)
        dinv = {}
        for lane in range(32):
            k = 1 << ((int(flags[lane]) >> 16) & 7)
            fw = int(d[lane, 2])                    # fetch word: slot | lanes-per-row log2 << 24 | blocks << 27
            assert (fw >> 24) & 7 == (int(flags[lane]) >> 16) & 7
            cs = [int(d[lane, 0]) & 0xFFFF, (int(d[lane, 0]) >> 16) & 0xFFFF, int(d[lane, 1]) & 0xFFFF, (int(d[lane, 1]) >> 16) & 0xFFFF]
            assert (fw >> 27) & 7 == sum(c != 0xFFFF for c in cs) and all(c != 0xFFFF for c in cs[:(fw >> 27) & 7])
            for t, c in enumerate(cs):
                if c == 0xFFFF:
                    continue
                slot = (fw & 0xFFFFFF) + t * k
                covered[slot] += 1
                nblk[lane] += 1
                if c == 0xFFFE:
                    assert is_u and t == 0 and (flags[lane] & (1 << 24))
                    dinv[lane] = lu[slot]
                    continue
                assert slot_col[slot] == c
                assert c not in this_level, 'gather inside the level that writes the value'
                if is_u:
                    assert left_u[c] == 0, 'gather of a value that is not yet final (backward)'
                else:
                    assert left_l[c] == 0 and not np.isnan(y[c]).any(), 'gather of a value that is not yet final (forward)'
                part[lane] += lu[slot] @ y[c]
        lane = 0
        while lane < 32:
            f = int(flags[lane])
            row = f & 0xFFFF
            span = (f >> 19) & 31
            if not f & (1 << 24):
                assert row == 0xFFFF and span == 0 and (part[lane] == 0).all()
                lane += 1
                continue
            k = 1 << ((f >> 16) & 7)
            assert span == k - 1 and lane % k == 0 and k <= (1 << rounds)
            for j in range(1, k):
                assert ((int(flags[lane + j]) >> 19) & 31) == k - 1 - j and not flags[lane + j] & (1 << 24)
                assert ((int(flags[lane + j]) >> 16) & 7) == ((f >> 16) & 7)
            v = part[lane:lane + k].sum(axis=0)
            cont = f < 0                            # RL_CONT (sign bit): a later piece of a row too long for one wave
            assert cont == bool(started[row]), 'the first piece starts from rhs / y, the others continue from x[row]'
            if cont:
                assert flags[0] & (1 << 25), 'a continuation is ordered after the piece before it'
                npieces += 1
            b = y[row] if (is_u or cont) else rhs[row]
            acc = b - v
            left[row] -= int(nblk[lane:lane + k].sum())
            assert left[row] >= 0
            if is_u:
                assert (lane in dinv) == (left[row] == 0), 'the pivot comes with the last piece'
            y[row] = dinv[lane] @ acc if lane in dinv else acc
            started[row] = True
            this_level.add(row)
            lane += k
    assert (covered == 1).all()
    assert (left_l == 0).all() and (left_u == 0).all()
    if name == 'long_rows':
        assert npieces > 0
    assert np.allclose(y, x, rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize('name', ['case118', 'case2869pegase'])
def test_row_lane_cta_programs(name):
    """Per-warp programs of k_solve_rows_cta / k_solve_rows_deep (symbolic.cpp, rc_*): every wave belongs to
    exactly one warp, a warp's waves and segments are in order, the waves of one level share a segment,
    a segment never mixes levels unless it is a run of one-wave levels (which is warp 0's), and the
    per-warp record streams are copies of the waves' records followed by the two idle waves."""
    case = load_case(name)
    S = _lib.symbolic_arrays(case.adj_ptr, case.adj_idx, case.bus_types_code())
    desc = S['rl_desc'].reshape(-1, 32, 4)
    nw = len(desc) - 2
    W = 7
    lists = S['rc_lists']
    off, ent = lists[:W + 1], lists[W + 1:].reshape(-1, 2)
    rc_desc = S['rc_desc'].reshape(-1, 32, 4)
    rc_seg = S['rc_seg']
    nseg = len(S['rc_segend'])
    assert off[0] == 0 and off[W] == len(ent) and len(rc_desc) >= len(ent) + 6 and len(rc_seg) >= len(ent) + 6
    level = np.cumsum([(int(desc[w, 0, 3]) >> 25) & 1 for w in range(nw)])       # level index of every wave
    owner = -np.ones(nw, dtype=int)
    seg_of = -np.ones(nw, dtype=int)
    for q in range(W):
        e = ent[off[q]:off[q + 1]]
        assert (e[-2:, 0] == [nw, nw + 1]).all() and (e[-2:, 1] == nseg).all()
        assert (np.diff(e[:-2, 0]) > 0).all() and (np.diff(e[:, 1]) >= 0).all()
        for i, (wv, sg) in enumerate(e[:-2]):
            assert owner[wv] == -1
            owner[wv], seg_of[wv] = q, sg
        for i, (wv, sg) in enumerate(e):
            assert (rc_desc[off[q] + i] == desc[wv]).all() and rc_seg[off[q] + i] == sg
    assert (owner >= 0).all() and seg_of.max() == nseg - 1 and (np.diff(seg_of) >= 0).all()
    for sg in range(nseg):
        ws = np.nonzero(seg_of == sg)[0]
        lv = level[ws]
        if len(set(lv)) > 1:                        # a run of one-wave levels: one warp, in order
            assert len(set(lv)) == len(ws) and (owner[ws] == 0).all()
        else:                                       # one level: dealt round-robin
            assert (owner[ws] == np.arange(len(ws)) % W).all()
    # a level never spans two segments
    for lv in range(1, level.max() + 1):
        assert len(set(seg_of[level == lv])) == 1
    segend = S['rc_segend']
    assert (np.diff(segend) >= 0).all() and segend[-1] <= int(S['lptr'][-1]) + int(S['uptr'][-1]) + case.N


def _write_matpower(path, name, buses, branches, gens, base=100.0):
    def mat(a):
        return '\n'.join('\t' + '\t'.join(repr(float(v)) for v in row) + ';' for row in a)
    with open(path, 'w') as f:
        f.write('function mpc = %s\n%% test case\nmpc.version = \'2\';\nmpc.baseMVA = %r;\n' % (name, base))
        f.write('%% bus data\nmpc.bus = [\n%s\n];\n' % mat(buses))
        f.write('mpc.gen = [\n%s\n];\n' % mat(gens))
        f.write('mpc.branch = [ %% fbus tbus r x b ...\n%s\n];\n' % mat(branches))


@pytest.mark.parametrize('name', ['case9', 'case118'])
def test_matpower_reader(name, tmp_path):
    """A MATPOWER .m case loads to the same CaseData as the reference-format tables, also from a
    file on another MVA base, and converts to the reference's .xlsx layout."""
    from common import load_tables
    from helmpy_b200.core.matpower import create_case_data_object_from_matpower, matpower_to_xlsx
    from helmpy_b200.core.classes import create_case_data_object_from_xlsx
    buses, branches, gens = load_tables(name)
    ref = load_case(name)
    m = str(tmp_path / (name + '.m'))
    _write_matpower(m, name, buses, branches, gens)
    c = create_case_data_object_from_matpower(m)
    assert c.name == name and c.N == ref.N and c.slack == ref.slack
    for attr in ('Pd', 'Qd', 'Pg', 'Qgmax', 'Qgmin', 'Yshunt', 'Ytrans_val', 'Y_val', 'adj_idx', 'phase_val'):
        assert np.array_equal(getattr(c, attr), getattr(ref, attr)), attr
    # the same grid written on a 50 MVA base (z halves, line charging doubles)
    b2 = branches.copy()
    b2[:, 2:4] *= 0.5
    b2[:, 4] *= 2.0
    m2 = str(tmp_path / (name + '_50.m'))
    _write_matpower(m2, name, buses, b2, gens, base=50.0)
    c2 = create_case_data_object_from_matpower(m2)
    assert np.allclose(c2.Ytrans_val, ref.Ytrans_val, rtol=1e-14, atol=0) and np.allclose(c2.Y_val, ref.Y_val, rtol=1e-13)
    x = str(tmp_path / (name + '.xlsx'))
    matpower_to_xlsx(m, x)
    c3 = create_case_data_object_from_xlsx(x)
    assert np.array_equal(c3.Ytrans_val, ref.Ytrans_val) and np.array_equal(c3.Pd, ref.Pd)


def test_bench_roofline_bytes_match_the_per_order_table():
    """bench.roofline_model's closed forms equal a brute-force walk over every order of every pass with
    the per-bus byte counts of DESIGN.md section 5."""
    import importlib.util
    spec = importlib.util.spec_from_file_location('bench', os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'bench.py'))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    N, n_pv, nslots = 1000, 150, 6000
    ncoef = np.array([[23, 25, 29, 0], [17, 0, 0, 0], [5, 7, 0, 0]], dtype=np.int32)
    stats = {'ms_kernel': {'solve': 1.0, 'rhs_conv': 1.0, 'eps': 1.0, 'factor': 1.0}, 'steps': 10}
    for shared in (True, False):
        k = bench.roofline_model({'n_slots': nslots, 'n_pv': n_pv}, ncoef, stats, N, shared_first_pass=shared)
        solve = conv = eps = fac = conv8 = eps8 = 0.0
        for row in ncoef:
            for pidx, m in enumerate(row[row > 0]):
                private = not (shared and pidx == 0)        # first passes run on ONE shared factorization (12 nnz / R)
                if private:
                    fac += (64 + 32) * nslots
                for j in range(1, int(m)):
                    solve += (32 * nslots if private else 0) + 32 * N
                    i = j & 3
                    pq = (32 * j + (48 if j >= 4 else 0)) if (j < 4 or i == 0) else (16 + 16 * (4 * i - 4))
                    conv += (pq + 64) * (N - n_pv) + (32 * j + 64) * n_pv
                    conv8 += 64 * N                          # x in, rhs out, V and W columns written once
                    if j % 2 == 0:
                        eps += (32 * j + 64) * N
                        eps8 += (16 * (j + 1) + 16) * N      # 16 m N read + 16 N write, m = j + 1 terms
        assert k['solve']['bytes'] == solve and k['factor']['bytes'] == fac
        assert k['rhs_conv']['bytes_8d'] == conv8 and k['eps']['bytes_8d'] == eps8
        assert abs(k['rhs_conv']['bytes'] - conv) < 1e-6 * conv and abs(k['eps']['bytes'] - eps) < 1e-6 * eps
