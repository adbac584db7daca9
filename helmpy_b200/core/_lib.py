"""ctypes binding of libhelm_b200.so (the C ABI declared in include/helm_b200.h).

There is no CPU fallback: if the CUDA library is missing this module raises at
import of the solver entry points, and if no CUDA device is usable the C ABI
returns HELM_E_CUDA, which is raised as ``HelmB200Error``.
"""

import ctypes as C
import os

import numpy as np

HELM_MAX_PASSES = 16
ST_RUNNING, ST_CONVERGED, ST_DIVERGED, ST_SINGULAR, ST_PASS_LIMIT = 0, 2, 3, 4, 5

_LIB_PATH = os.environ.get("HELM_B200_LIB") or os.path.join(
    os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "libhelm_b200.so")


class HelmB200Error(RuntimeError):
    pass


class GridHost(C.Structure):
    _fields_ = [
        ("n_bus", C.c_int32), ("slack", C.c_int32),
        ("adj_ptr", C.c_void_p), ("adj_idx", C.c_void_p),
        ("ytrans", C.c_void_p), ("yfull", C.c_void_p), ("yshunt", C.c_void_p),
        ("phase_ptr", C.c_void_p), ("phase_idx", C.c_void_p), ("phase_val", C.c_void_p),
        ("vsp", C.c_void_p), ("pg", C.c_void_p), ("pd", C.c_void_p), ("qd", C.c_void_p),
        ("qgmax", C.c_void_p), ("qgmin", C.c_void_p), ("bus_type", C.c_void_p),
        ("n_branch", C.c_int32), ("br_from", C.c_void_p), ("br_to", C.c_void_p), ("br_y", C.c_void_p),
        ("br_sh_from", C.c_void_p), ("br_sh_to", C.c_void_p), ("br_ph_ft", C.c_void_p),
        ("br_ph_tf", C.c_void_p),
    ]


class Params(C.Structure):
    _fields_ = [
        ("mismatch", C.c_double), ("max_coefficients", C.c_int32),
        ("enforce_q_limits", C.c_int32), ("group", C.c_int32),
        ("use_graph", C.c_int32), ("poll_every", C.c_int32), ("max_resident", C.c_int32),
        ("pv_bus_model", C.c_int32), ("dsb_method", C.c_int32), ("dsb_model", C.c_int32),
        ("record_switches", C.c_int32),
    ]


class PlanInfo(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "n_bus", "nnz_adj", "n_l_blocks", "n_u_blocks", "n_slots", "n_l_levels",
        "n_u_levels", "n_factor_updates", "n_pv", "reserved")]


class Stats(C.Structure):
    _fields_ = [
        ("steps", C.c_int64), ("kernel_launches", C.c_int64), ("graph_launches", C.c_int64),
        ("ms_total", C.c_double), ("ms_kernel", C.c_double * 12),
    ]


KERNEL_NAMES = ("assemble", "factor", "solve", "rhs_conv", "advance", "qlimit", "pass_end", "init", "eps", "fcache")

EXPORTED_SYMBOLS = (
    "helm_last_error", "helm_plan_create", "helm_plan_destroy", "helm_plan_get_info",
    "helm_plan_get_perm", "helm_plan_set_outages", "helm_workspace_bytes", "helm_solve_batch_device",
    "helm_solve_batch_host", "helm_nr_batch_host", "helm_last_state_host", "helm_last_switches_host", "helm_last_series_host", "helm_last_ploss_host", "helm_debug_factor_solve", "helm_debug_solve_time", "helm_debug_timeline", "helm_measure_peaks", "helm_measure_latencies",
    "helm_symbolic_create", "helm_symbolic_destroy", "helm_symbolic_array",
)

_lib = None


def lib_path():
    return _LIB_PATH


def load():
    """Load the shared library; raise loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise HelmB200Error(
            "CUDA library %s not found: build it with helmpy_b200/csrc/build.sh "
            "(or `python -c 'import __graft_entry__ as g; g.build()'`). "
            "helmpy_b200 has no CPU fallback." % _LIB_PATH)
    lib = C.CDLL(_LIB_PATH)
    lib.helm_last_error.restype = C.c_char_p
    lib.helm_plan_create.argtypes = [C.POINTER(GridHost), C.POINTER(C.c_void_p)]
    lib.helm_plan_destroy.argtypes = [C.c_void_p]
    lib.helm_plan_destroy.restype = None
    lib.helm_plan_get_info.argtypes = [C.c_void_p, C.POINTER(PlanInfo)]
    lib.helm_plan_get_perm.argtypes = [C.c_void_p, C.c_void_p]
    lib.helm_plan_set_outages.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
    lib.helm_workspace_bytes.argtypes = [C.c_void_p, C.c_int32, C.POINTER(Params)]
    lib.helm_workspace_bytes.restype = C.c_size_t
    lib.helm_solve_batch_device.argtypes = [
        C.c_void_p, C.c_int32, C.c_void_p, C.POINTER(Params), C.c_void_p, C.c_size_t,
        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(Stats)]
    lib.helm_solve_batch_host.argtypes = [
        C.c_void_p, C.c_int32, C.c_void_p, C.POINTER(Params), C.c_void_p, C.c_void_p,
        C.c_void_p, C.c_void_p, C.POINTER(Stats)]
    lib.helm_nr_batch_host.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_double, C.c_int32, C.c_int32,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.helm_last_state_host.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
    lib.helm_last_switches_host.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.helm_last_series_host.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    lib.helm_last_ploss_host.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
    lib.helm_debug_factor_solve.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p,
                                            C.c_void_p, C.c_void_p]
    lib.helm_debug_solve_time.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_double)]
    lib.helm_debug_timeline.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
    lib.helm_measure_peaks.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.helm_measure_latencies.argtypes = [C.c_void_p]
    lib.helm_symbolic_create.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.POINTER(C.c_void_p)]
    lib.helm_symbolic_destroy.argtypes = [C.c_void_p]
    lib.helm_symbolic_destroy.restype = None
    lib.helm_symbolic_array.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p),
                                        C.POINTER(C.c_int32)]
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().helm_last_error()
        raise HelmB200Error("%s failed (code %d): %s" % (what, rc, (msg or b"").decode()))


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def symbolic_arrays(adj_ptr, adj_idx, bus_type=None):
    """Host-only symbolic analysis -> dict of int32 numpy arrays (no CUDA)."""
    lib = load()
    adj_ptr = np.ascontiguousarray(adj_ptr, dtype=np.int32)
    adj_idx = np.ascontiguousarray(adj_idx, dtype=np.int32)
    bt = None if bus_type is None else np.ascontiguousarray(bus_type, dtype=np.uint8)
    h = C.c_void_p()
    rc = lib.helm_symbolic_create(len(adj_ptr) - 1, _ptr(adj_ptr), _ptr(adj_idx),
                                  None if bt is None else _ptr(bt), C.byref(h))
    if rc != 0:
        raise HelmB200Error("helm_symbolic_create failed (code %d)" % rc)
    out = {}
    try:
        for name in ("perm", "iperm", "adj_ptr", "adj_idx", "adj_src", "adj_slot", "lptr", "lcol",
                     "llev_ptr", "urow", "qof", "ulev_ptr", "uptr", "ucol", "dslot", "fptr",
                     "fsrc", "fdst", "fdst_local", "fchunk_row", "fchunk_lpr", "flev_chunk", "frow_off",
                     "slot_src", "slot_row", "slot_col", "wave_tab", "smeta", "scm", "solve_chunks", "solve_segs", "one_lists",
                     "rl_desc", "rc_lists", "rc_segend", "rc_desc", "rc_seg"):
            data = C.c_void_p()
            n = C.c_int32()
            rc = lib.helm_symbolic_array(h, name.encode(), C.byref(data), C.byref(n))
            if rc != 0:
                raise HelmB200Error("helm_symbolic_array(%s) failed" % name)
            if n.value:
                buf = (C.c_int32 * n.value).from_address(data.value)
                out[name] = np.frombuffer(buf, dtype=np.int32).copy()
            else:
                out[name] = np.zeros(0, dtype=np.int32)
    finally:
        lib.helm_symbolic_destroy(h)
    return out


def _result_array(n_scen, n_bus):
    """complex128 [n_scen, n_bus] host array for the voltages.  Large results are placed in
    page-locked memory from torch's caching host allocator (torch is the plumbing for memory here):
    the device-to-host copy of helm_solve_batch_host then runs at PCIe speed instead of through a
    pageable staging buffer, and the block is recycled when the array is dropped.  Every element is
    written by the C call (diverged scenarios get zeros)."""
    if n_scen * n_bus * 16 >= (8 << 20):
        try:
            import torch
            if torch.cuda.is_available():
                return torch.empty((n_scen, n_bus), dtype=torch.complex128, pin_memory=True).numpy()
        except Exception:
            pass
    return np.zeros((n_scen, n_bus), dtype=np.complex128)


class Plan:
    """Owns a ``helm_plan``: symbolic analysis + device-resident grid for one case."""

    def __init__(self, case):
        lib = load()
        self._keep = k = {}
        k["adj_ptr"] = np.ascontiguousarray(case.adj_ptr, dtype=np.int32)
        k["adj_idx"] = np.ascontiguousarray(case.adj_idx, dtype=np.int32)
        k["ytrans"] = np.ascontiguousarray(case.Ytrans_val, dtype=np.complex128)
        k["yfull"] = np.ascontiguousarray(case.Y_val, dtype=np.complex128)
        k["yshunt"] = np.ascontiguousarray(case.Yshunt, dtype=np.complex128)
        k["phase_ptr"] = np.ascontiguousarray(case.phase_ptr, dtype=np.int32)
        k["phase_idx"] = np.ascontiguousarray(case.phase_idx, dtype=np.int32)
        k["phase_val"] = np.ascontiguousarray(case.phase_val, dtype=np.complex128)
        for name, attr in (("vsp", "V"), ("pg", "Pg"), ("pd", "Pd"), ("qd", "Qd"),
                           ("qgmax", "Qgmax"), ("qgmin", "Qgmin")):
            k[name] = np.ascontiguousarray(getattr(case, attr), dtype=np.float64)
        k["bus_type"] = np.ascontiguousarray(case.bus_types_code(), dtype=np.uint8)
        g = GridHost()
        g.n_bus = case.N
        g.slack = int(case.slack)
        for name in ("adj_ptr", "adj_idx", "ytrans", "yfull", "yshunt", "phase_ptr", "phase_idx",
                     "phase_val", "vsp", "pg", "pd", "qd", "qgmax", "qgmin", "bus_type"):
            setattr(g, name, k[name].ctypes.data)
        bc = getattr(case, 'branch_contrib', None)
        g.n_branch = 0
        if bc is not None and len(bc['f']):
            k["br_from"] = np.ascontiguousarray(bc['f'], dtype=np.int32)
            k["br_to"] = np.ascontiguousarray(bc['t'], dtype=np.int32)
            for name, key in (("br_y", 'y'), ("br_sh_from", 'sh_f'), ("br_sh_to", 'sh_t'),
                              ("br_ph_ft", 'ph_ft'), ("br_ph_tf", 'ph_tf')):
                k[name] = np.ascontiguousarray(bc[key], dtype=np.complex128)
            g.n_branch = len(k["br_from"])
            for name in ("br_from", "br_to", "br_y", "br_sh_from", "br_sh_to", "br_ph_ft", "br_ph_tf"):
                setattr(g, name, k[name].ctypes.data)
        self.n_branch = int(g.n_branch)
        self.handle = C.c_void_p()
        check(lib.helm_plan_create(C.byref(g), C.byref(self.handle)), "helm_plan_create")
        self.N = case.N
        info = PlanInfo()
        check(lib.helm_plan_get_info(self.handle, C.byref(info)), "helm_plan_get_info")
        self.info = {n: getattr(info, n) for n, _ in PlanInfo._fields_}

    def __del__(self):
        try:
            if getattr(self, "handle", None) and self.handle.value and _lib is not None:
                _lib.helm_plan_destroy(self.handle)
                self.handle = C.c_void_p()
        except Exception:
            pass

    def perm(self):
        out = np.empty(self.N, dtype=np.int32)
        check(load().helm_plan_get_perm(self.handle, _ptr(out)), "helm_plan_get_perm")
        return out

    @staticmethod
    def make_params(mismatch, max_coefficients, enforce_q_limits, group=0, use_graph=True,
                    poll_every=0, max_resident=0, pv_bus_model=2, dsb_method=0, dsb_model=False,
                    record_switches=False):
        p = Params()
        p.mismatch = float(mismatch)
        p.max_coefficients = int(max_coefficients)
        p.enforce_q_limits = int(bool(enforce_q_limits))
        p.group = int(group)
        p.use_graph = int(bool(use_graph))
        p.poll_every = int(poll_every)
        p.max_resident = int(max_resident)
        p.pv_bus_model = int(pv_bus_model)
        p.dsb_method = int(dsb_method or 0)
        p.dsb_model = int(bool(dsb_model))
        p.record_switches = int(bool(record_switches))
        return p

    def workspace_bytes(self, n_scen, params):
        return int(load().helm_workspace_bytes(self.handle, n_scen, C.byref(params)))

    def solve_host(self, scales, params, outages=None):
        """Host-buffer call: numpy in, numpy out (H2D/D2H inside).  ``outages[b]`` = index of the
        branch removed in scenario b (-1: none)."""
        lib = load()
        scales = np.ascontiguousarray(scales, dtype=np.float64)
        B = len(scales)
        if outages is not None:
            outages = np.ascontiguousarray(outages, dtype=np.int32)
            if len(outages) != B:
                raise HelmB200Error("outages and scales must have the same length")
            check(lib.helm_plan_set_outages(self.handle, B, _ptr(outages)), "helm_plan_set_outages")
        vout = _result_array(B, self.N)
        status = np.zeros(B, dtype=np.int32)
        ncoef = np.zeros((B, HELM_MAX_PASSES), dtype=np.int32)
        nswitch = np.zeros(B, dtype=np.int32)
        stats = Stats()
        check(lib.helm_solve_batch_host(self.handle, B, _ptr(scales), C.byref(params), _ptr(vout),
                                        _ptr(status), _ptr(ncoef), _ptr(nswitch), C.byref(stats)),
              "helm_solve_batch_host")
        return vout, status, ncoef, nswitch, stats_dict(stats)

    def nr_host(self, scales, mismatch, max_iterations, enforce_q_limits):
        """Batched Newton-Raphson cross-check on the GPU: numpy in, numpy out."""
        lib = load()
        scales = np.ascontiguousarray(scales, dtype=np.float64)
        B = len(scales)
        vout = _result_array(B, self.N)
        status = np.zeros(B, dtype=np.int32)
        iters = np.zeros((B, HELM_MAX_PASSES), dtype=np.int32)
        nswitch = np.zeros(B, dtype=np.int32)
        check(lib.helm_nr_batch_host(self.handle, B, _ptr(scales), float(mismatch), int(max_iterations),
                                     int(bool(enforce_q_limits)), _ptr(vout), _ptr(status), _ptr(iters),
                                     _ptr(nswitch)), "helm_nr_batch_host")
        return vout, status, iters, nswitch

    def solve_device(self, d_scale, params, workspace, d_vout, d_status, d_ncoef, d_nswitch,
                     stream_ptr):
        """Device-pointer call on torch tensors (resident inputs)."""
        lib = load()
        stats = Stats()
        check(lib.helm_solve_batch_device(
            self.handle, d_scale.numel(), d_scale.data_ptr(), C.byref(params),
            workspace.data_ptr(), workspace.numel() * workspace.element_size(),
            d_vout.data_ptr(), d_status.data_ptr(), d_ncoef.data_ptr(), d_nswitch.data_ptr(),
            stream_ptr, C.byref(stats)), "helm_solve_batch_device")
        return stats_dict(stats)

    def last_state(self, n_scen):
        """(Qg [B, N], bus types [B, N]) at the end of the last ``solve_host`` of n_scen scenarios."""
        qg = np.zeros((n_scen, self.N), dtype=np.float64)
        bt = np.zeros((n_scen, self.N), dtype=np.uint8)
        check(load().helm_last_state_host(self.handle, n_scen, _ptr(qg), _ptr(bt)), "helm_last_state_host")
        return qg, bt

    def last_series(self, n_scen, scen, n_terms):
        """(V [n_terms, N], H [n_terms, N]) series of the last pass of scenario ``scen`` after a
        ``solve_host`` of n_scen resident scenarios: H = W on PQ buses, barras_CC on PV buses."""
        V = np.zeros((n_terms, self.N), dtype=np.complex128)
        H = np.zeros((n_terms, self.N), dtype=np.complex128)
        check(load().helm_last_series_host(self.handle, n_scen, scen, n_terms, _ptr(V), _ptr(H)),
              "helm_last_series_host")
        return V, H

    def last_switches(self, n_scen, nswitch):
        """PV->PQ switches of the last ``solve_host`` made with record_switches: per scenario a list
        of (pass, bus, violating Qg) sorted by pass and bus (the order the reference visits them,
        helm.py:479)."""
        cap = max(1, int(np.max(nswitch)))
        bus = np.zeros((n_scen, cap), dtype=np.int32)
        pas = np.zeros((n_scen, cap), dtype=np.int32)
        q = np.zeros((n_scen, cap), dtype=np.float64)
        check(load().helm_last_switches_host(self.handle, n_scen, cap, _ptr(bus), _ptr(pas), _ptr(q)),
              "helm_last_switches_host")
        return [sorted((int(pas[b, k]), int(bus[b, k]), float(q[b, k])) for k in range(cap) if bus[b, k] >= 0)
                for b in range(n_scen)]

    def last_ploss(self, n_scen, max_coefficients, dsb_method, pv_bus_model):
        """Ploss series [B, max_coefficients+1] of the last distributed-slack ``solve_host``."""
        out = np.zeros((n_scen, max_coefficients + 1), dtype=np.float64)
        check(load().helm_last_ploss_host(self.handle, n_scen, int(dsb_method), int(pv_bus_model), _ptr(out)),
              "helm_last_ploss_host")
        return out

    def debug_solve_time(self, kernel=0, ablation=0, reps=50):
        """Device time (ms) of one triangular solve of one scenario, one kernel alone on the GPU."""
        ms = C.c_double()
        check(load().helm_debug_solve_time(self.handle, kernel, ablation, reps, C.byref(ms)), "helm_debug_solve_time")
        return ms.value

    def debug_factor_solve(self, bus_types, rhs, group=1):
        bus_types = np.ascontiguousarray(bus_types, dtype=np.uint8)
        rhs = np.ascontiguousarray(rhs, dtype=np.float64)
        B = bus_types.shape[0]
        x = np.zeros_like(rhs)
        check(load().helm_debug_factor_solve(self.handle, B, group, _ptr(bus_types), _ptr(rhs),
                                             _ptr(x)), "helm_debug_factor_solve")
        return x


def stats_dict(stats):
    return {
        "steps": int(stats.steps), "kernel_launches": int(stats.kernel_launches),
        "graph_launches": int(stats.graph_launches), "ms_total": float(stats.ms_total),
        "ms_kernel": {KERNEL_NAMES[i]: float(stats.ms_kernel[i]) for i in range(len(KERNEL_NAMES))},
        # factor cache: pass starts that found their factors already in the pool / factorizations performed
        "factor_cache_hits": int(stats.ms_kernel[10]), "factorizations": int(stats.ms_kernel[11]),
    }


def measure_peaks():
    a, b = C.c_double(), C.c_double()
    check(load().helm_measure_peaks(C.byref(a), C.byref(b)), "helm_measure_peaks")
    return {"dfma_tflops": a.value, "copy_gbs": b.value}


def measure_latencies():
    """Dependent-operation latencies (cycles): DADD, DFMA, fp64 shuffle + add, LDS, L2-hit load."""
    out = np.zeros(5)
    check(load().helm_measure_latencies(_ptr(out)), "helm_measure_latencies")
    return dict(zip(("dadd", "dfma", "shfl_dadd", "lds", "l2_load"), out.tolist()))
