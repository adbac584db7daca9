#!/usr/bin/env python
"""Benchmark of the HELM + Pade hot path (contract: see the task prompt / DESIGN.md §7).

    python bench.py --gpus N --steps K --warmup W            # this build (CUDA)
    python bench.py --impl reference --gpus N --steps K --warmup W   # CPU arm

A *step* is one pass of the hot path over one batch of synthetic input: B
load-scaling scenarios of case2869pegase per GPU (scales ~ U[0.80, 1.05],
numpy default_rng(seed)), each solved exactly as helm(case, mismatch=1e-8,
scale=s, max_coefficients=40) would (Q limits on).  Metric: converged power-flow
cases per second, whole job (all ranks).  `value` is measured with the scenario
scales and the grid already resident in HBM; `e2e` goes through the host-buffer
C-ABI call (`helm_batch`): scales host->device and voltages device->host inside
the timed region.  4096 scenarios are resident at a time (7 MB each, 28 GB:
far beyond the 126 MB L2, so every step streams from HBM); the rest of the batch
queues on the device and takes over slots as scenarios finish.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, 'tests'))

CASE = 'case2869pegase'
MISMATCH = 1e-8
MAX_COEF = 40
SCALE_LO, SCALE_HI = 0.80, 1.05
RESIDENT = [4096]
METRIC = 'converged_power_flow_cases_per_sec'
UNIT = 'cases/s'


def scenario_scales(n, seed):
    return np.random.default_rng(seed).uniform(SCALE_LO, SCALE_HI, n)


# ------------------------------------------------------------------------------------------
# CPU arm: the oracle port (the reference is pure Python and cannot travel to the GPU box;
# oracle/helm_oracle.py restates its algorithm and is pinned to it, see tests/test_oracle.py)
# ------------------------------------------------------------------------------------------
def _cpu_worker(args):
    os.environ['OPENBLAS_NUM_THREADS'] = '1'
    os.environ['OMP_NUM_THREADS'] = '1'
    from common import load_case
    from oracle import helm_oracle as ho
    scales = args
    case = load_case(CASE)
    nconv = 0
    for s in scales:
        V, _ = ho.helm_oracle(case, mismatch=MISMATCH, scale=float(s), max_coefficients=MAX_COEF)
        nconv += V is not None
    return nconv


def cpu_cases_per_sec(n_scen, cores, seed=0):
    """Solve n_scen scenarios with the oracle on `cores` processes; return (cases/s, elapsed)."""
    import multiprocessing as mp
    scales = scenario_scales(n_scen, seed)
    chunks = [c for c in np.array_split(scales, cores) if len(c)]
    ctx = mp.get_context('spawn')
    with ctx.Pool(len(chunks)) as pool:
        pool.map(_cpu_worker, [c[:1] for c in chunks])          # warm-up: imports, case load
        t0 = time.perf_counter()
        nconv = sum(pool.map(_cpu_worker, chunks))
        dt = time.perf_counter() - t0
    return nconv / dt, dt, nconv


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per_step = 4 * cores          # ~3-4 s of the oracle port per step on every core
    vals = []
    for it in range(args.warmup + args.steps):
        v, dt, nconv = cpu_cases_per_sec(per_step, cores, seed=it)
        if it >= args.warmup:
            vals.append((v, dt))
    value = float(np.mean([v for v, _ in vals]))
    ms = float(np.mean([dt for _, dt in vals])) * 1e3
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': workload_config(args.batch, args.gpus),      # the CUDA arm's workload; each step times a bounded sample of it
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                         'sample': '%d scenarios per step over %d processes (numpy/scipy oracle port of '
                                   'the reference algorithm; the reference itself is pure Python and is '
                                   '~12x slower per case, DESIGN.md §7)' % (per_step, cores)},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(batch_per_gpu, n_gpus):
    return {
        'workload': '%s load-scaling scenarios, HELM PV2 + Pade, Q limits on, mismatch 1e-8, '
                    'max_coefficients 40, scale~U[%.2f,%.2f]' % (CASE, SCALE_LO, SCALE_HI),
        'grid': CASE, 'n_bus': 2869, 'scenarios_per_gpu': batch_per_gpu,
        'scenarios_total': batch_per_gpu * n_gpus, 'parallelism': 'scenario-sharded x%d' % n_gpus,
        'l2': 'working set >> 126 MB L2 (7 MB per resident scenario), no flush needed',
        'resident_scenarios_per_gpu': RESIDENT[0],
    }


# ------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.stop_flag = False
        self.proc = None

    def run(self):
        q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.index), '--query-gpu=' + q, '--format=csv,noheader,nounits',
                 '-lms', '200'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(',')])
                if self.stop_flag:
                    break
        except Exception:
            pass

    def stop(self):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith('active'):
                        reasons.add(n)
            except Exception:
                continue
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        return {'sm_mhz': float(np.median(sm)), 'sm_max_mhz': float(max(mx)), 'reasons': sorted(reasons),
                'samples': len(sm)}


def load_ncu_traffic(kernel):
    """DRAM bytes of one launch of `kernel` from the committed `ncu --set full` capture
    (profiles/ncu_traffic.json; dram__bytes_read.sum + dram__bytes_write.sum) or None."""
    p = os.path.join(REPO, 'profiles', 'ncu_traffic.json')
    try:
        return json.load(open(p)).get(kernel)
    except Exception:
        return None


def load_peaks():
    p = os.path.join(REPO, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


def run_cuda(args):
    import torch
    import torch.distributed as dist
    from common import load_case
    import helmpy_b200
    from helmpy_b200.core import _lib
    from helmpy_b200.core.helm import get_plan

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        if os.environ.get('NCCL_DEBUG', 'VERSION').upper() == 'VERSION':
            os.environ['NCCL_DEBUG'] = 'WARN'
        os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')   # keep stdout to the one JSON line
        dist.init_process_group('nccl', device_id=dev)
    B = args.batch
    case = load_case(CASE)
    N = case.N
    plan = get_plan(case)
    prm = _lib.Plan.make_params(MISMATCH, MAX_COEF, True, args.group, True, 0, args.resident)
    ws = torch.empty(plan.workspace_bytes(B, prm), dtype=torch.uint8, device=dev)
    d_vout = torch.zeros(B, N, dtype=torch.complex128, device=dev)
    d_status = torch.zeros(B, dtype=torch.int32, device=dev)
    d_ncoef = torch.zeros(B, _lib.HELM_MAX_PASSES, dtype=torch.int32, device=dev)
    d_nsw = torch.zeros(B, dtype=torch.int32, device=dev)
    gathered = None
    if world > 1:
        gathered = torch.empty(world, B, N, 2, dtype=torch.float64, device=dev)
    stream = torch.cuda.Stream(device=dev)
    total_steps = args.warmup + args.steps
    # every step gets fresh scenarios (rank- and step-dependent seed), resident in HBM beforehand
    d_scales = [torch.from_numpy(scenario_scales(B, 1000 * rank + it)).to(dev) for it in range(total_steps)]
    torch.cuda.synchronize()

    launches = 0
    solves = 0
    conv_total = 0
    kernel_ms = 0.0

    def one_step(it):
        nonlocal launches, solves, conv_total, kernel_ms
        with torch.cuda.stream(stream):
            st = plan.solve_device(d_scales[it], prm, ws, d_vout, d_status, d_ncoef, d_nsw, stream.cuda_stream)
            if world > 1:   # the path's only exchange: final gather of the voltages (NCCL over NVLink)
                dist.all_gather_into_tensor(gathered.view(world * B, N, 2), torch.view_as_real(d_vout))
        return st

    def barrier():
        stream.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for it in range(args.warmup):
        one_step(it)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for it in range(args.warmup, total_steps):
        st = one_step(it)
        launches += st['kernel_launches']
        kernel_ms += st['ms_total']
        conv_total += int((d_status == _lib.ST_CONVERGED).sum().item())
        solves += int(torch.clamp(d_ncoef - 1, min=0).sum().item())
    ev1.record(stream)
    barrier()
    elapsed_ms = ev0.elapsed_time(ev1)
    if rank == 0:
        sampler.stop()
    t = torch.tensor([elapsed_ms, float(conv_total), float(launches), float(solves)], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        elapsed_ms = float(tmax[0].item())
        conv_all, launches_all, solves_all = float(tsum[1].item()), float(tsum[2].item()), float(tsum[3].item())
    else:
        conv_all, launches_all, solves_all = float(conv_total), float(launches), float(solves)
    value = conv_all / (elapsed_ms * 1e-3)

    # ---- e2e: host buffers through the public API, H2D + D2H inside the timed region -------
    host_scales = [scenario_scales(B, 5000 + 1000 * rank + it) for it in range(2)]
    res = None
    for it in range(2):     # warm-up: plan workspace, and the two page-locked result blocks a `res = helm_batch(...)` loop alternates between
        res = helmpy_b200.helm_batch(case, host_scales[it], mismatch=MISMATCH, max_coefficients=MAX_COEF, group=args.group,
                                     max_resident=args.resident)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    e2e_conv = 0
    e2e_steps = max(1, min(args.steps, 2))
    for it in range(e2e_steps):
        res = helmpy_b200.helm_batch(case, host_scales[it % 2], mismatch=MISMATCH, max_coefficients=MAX_COEF,
                                     group=args.group, max_resident=args.resident)
        e2e_conv += int(res.converged.sum())
    torch.cuda.synchronize()
    e2e_dt = time.perf_counter() - t0
    te = torch.tensor([e2e_dt, float(e2e_conv)], dtype=torch.float64, device=dev)
    if world > 1:
        tm = te.clone()
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        tsu = te.clone()
        dist.all_reduce(tsu, op=dist.ReduceOp.SUM)
        e2e_dt, e2e_conv = float(tm[0].item()), float(tsu[1].item())
    e2e_value = e2e_conv / e2e_dt
    h2d = 8 * B
    d2h = B * N * 16 + B * 4 * (2 + _lib.HELM_MAX_PASSES)

    # ---- per-kernel roofline: one extra profiled step (events around every kernel, no graph) ----
    roof = None
    kernels = None
    if rank == 0:
        os.environ['HELM_B200_PROFILE'] = '1'
        with torch.cuda.stream(stream):
            stp = plan.solve_device(d_scales[-1], prm, ws, d_vout, d_status, d_ncoef, d_nsw, stream.cuda_stream)
        stream.synchronize()
        os.environ.pop('HELM_B200_PROFILE', None)
        nc = d_ncoef.cpu().numpy()
        info = plan.info
        kernels = roofline_model(info, nc, stp, N)
        peak, how = load_peaks()
        for k in kernels.values():
            k['frac_of_hbm_peak'] = k['gbs'] / peak
        # dominant kernel = the largest of the main branch (solve -> eps -> rhs_conv is the critical
        # path of a step; the factorization runs beside it in the side branch, two steps deep)
        dom = max((kernels[k] for k in ('solve', 'rhs_conv', 'eps')), key=lambda k: k['ms'])
        tr = load_ncu_traffic(dom['name'])
        roof = {'bound': 'hbm', 'kernel': dom['name'], 'achieved': dom['gbs'], 'peak': peak, 'unit': 'GB/s',
                'frac': dom['gbs'] / peak, 'traffic': tr['dram_bytes'] if tr else None,
                'traffic_source': tr, 'peak_source': how,
                'share_of_step': dom['ms'] / max(1e-9, sum(kernels[k]['ms'] for k in ('solve', 'rhs_conv', 'eps'))),
                'share_of_step_note': 'share of the main-branch kernel time (solve + eps + rhs_conv); factorization overlaps in the side branch',
                'launches': dom['launches'], 'avg_launch_ms': dom['ms'] / max(1, dom['launches']),
                'algorithmic_bytes_per_launch': dom['bytes'] / max(1, dom['launches'])}
    if world > 1:
        dist.barrier()

    single_ms = None
    if rank == 0:
        # BASELINE config[1]: one case2869pegase case to 1e-8 (per-order latency path, CUDA graph)
        ts = []
        for _ in range(5):
            r1 = helmpy_b200.helm_batch(case, [1.0], mismatch=MISMATCH, max_coefficients=MAX_COEF)
            ts.append(r1.stats['ms_total'])
        single_ms = float(np.median(ts))
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            n = 12 * cores                # bounded sample: ~10 s of wall time on all host cores
            v, dt, nconv = cpu_cases_per_sec(n, cores, seed=0)
            cpu = {'value': v, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                   'sample': '%d scenarios of the same workload (seed 0), %d processes, %.1f s; numpy/scipy '
                             'oracle port (oracle/helm_oracle.py)' % (n, cores, dt)}
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': elapsed_ms / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': workload_config(B, world),
            'clocks': sampler.summary(),
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                    'steps': e2e_steps},
            'gpu_launches': int(launches_all),
            'roofline': roof, 'roofline_kernels': kernels, 'cpu_baseline': cpu,
            'single_case_ms': single_ms, 'scenario_orders': int(solves_all), 'converged': int(conv_all),
            'group': prm.group, 'device_ms_sum_rank0': kernel_ms,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def roofline_model(info, ncoef, stats, N):
    """Algorithmic HBM bytes per kernel / measured kernel time.

    ncoef[b, p] = coefficients of scenario b in pass p.  A pass with m coefficients
    runs orders j = 1..m-1.
    """
    m = ncoef[ncoef > 0].astype(np.float64)
    orders = float(np.sum(m - 1))                       # scenario-orders = triangular solves
    passes = float(len(m))
    sum_j = float(np.sum((m - 1) * m / 2))              # sum over orders of j
    half = np.floor((m - 1) / 2)
    sum_j_even = float(np.sum(half * (half + 1)))       # sum over the even orders of j
    nslots = info['n_slots']
    out = {}
    ms = stats['ms_kernel']
    steps = max(1, stats['steps'])
    # K2 solve: LU blocks (32 B each) once + rhs read + x write (16 B each per bus)
    b_solve = orders * (32.0 * nslots + 32.0 * N)
    # K1 k_rhs_conv at order j, per bus (x read, V[j], H[j], rhs written: 64 B at every order):
    #   PV bus: histories V and CC, read j each (32 j).  PQ bus, blocked W recurrence (block 4):
    #   j < 4 or j % 4 == 0: histories V and W read j each (32 j), at a block start also three
    #   partial sums written (48); j % 4 == i > 0: one partial read (16) and 4 i - 4 history values (16 each).
    # K3 k_eps, even orders only: epsilon anti-diagonal read j-1, write j+1, + x, V[j-1], Pade value
    #   read and written: 16 B each.
    def conv_bytes_pq(mm):
        j = np.arange(1, int(mm))
        i = j & 3
        walk = (j < 4) | (i == 0)
        b = np.where(walk, 32.0 * j + np.where(j >= 4, 48.0, 0.0), 16.0 + 16.0 * (4 * i - 4))
        return float(np.sum(b + 64.0))
    orders_even = float(np.sum(half))
    n_pv = float(info.get('n_pv', 0))
    um, cnt = np.unique(m, return_counts=True)
    b_pq = float(sum(c * conv_bytes_pq(v) for v, c in zip(um, cnt)))
    b_rhs = b_pq * (N - n_pv) + (32.0 * sum_j + 64.0 * orders) * n_pv
    b_eps = (32.0 * sum_j_even + 64.0 * orders_even) * N
    # K2f factor: read A blocks + write LU blocks (32 B per slot each way), per pass
    b_fac = passes * (64.0 + 32.0) * nslots       # assembly (write) + factorization (read, write)
    knames = {'solve': 'k_solve_flat', 'rhs_conv': 'k_rhs_conv', 'eps': 'k_eps', 'factor': 'k_factor_coop'}
    for name, b in (('solve', b_solve), ('rhs_conv', b_rhs), ('eps', b_eps), ('factor', b_fac)):
        t = ms.get(name, 0.0)
        out[name] = {'name': knames[name], 'ms': t, 'bytes': b, 'gbs': (b / (t * 1e-3) / 1e9) if t > 0 else 0.0,
                     'launches': steps}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='cuda', choices=['cuda', 'reference'])
    ap.add_argument('--batch', type=int, default=int(os.environ.get('HELM_BENCH_BATCH', '32768')),
                    help='scenarios per GPU per step')
    ap.add_argument('--group', type=int, default=int(os.environ.get('HELM_BENCH_GROUP', '1')))
    ap.add_argument('--resident', type=int, default=int(os.environ.get('HELM_BENCH_RESIDENT', '4096')),
                    help='scenarios resident on the GPU at once; the rest of the batch queues (continuous batching)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    RESIDENT[0] = args.resident
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == '__main__':
    main()
