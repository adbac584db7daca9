"""The five configurations of BASELINE.json, end to end through the public API (GPU box).

    python tests/run_configs.py [1 2 3 4 5]                      # one GPU
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tests/run_configs.py 3 5

Configs 3-5 name MATPOWER grids the reference does not ship (case2383wp, case9241pegase,
case13659pegase): tests/synth.py builds stand-ins of the same size from the shipped grids
(SURVEY 8d).  Every config is checked on sampled scenarios against the oracle
(oracle/helm_oracle.py; this script lives under tests/ because only test code may use the oracle) and, for config 4, against the Newton-Raphson
cross-check.  One JSON line per config; rank 0 appends them to gpurun_out/configs.jsonl.
"""
import contextlib
import io
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, 'tests'))
from common import load_case, golden_results  # noqa: E402
import synth  # noqa: E402
import helmpy_b200  # noqa: E402
from helmpy_b200.core.parallel import helm_batch_sharded  # noqa: E402
from oracle import helm_oracle as ho  # noqa: E402

WORLD = int(os.environ.get('WORLD_SIZE', '1'))
RANK = int(os.environ.get('RANK', '0'))


def polar_err(a, b):
    return float(np.max(np.abs(np.abs(a) - np.abs(b)))), float(np.max(np.abs(np.angle(a) - np.angle(b))))


def solve(case, scales, outages=None, **kw):
    """(V, status, ncoef, nswitch, seconds) over all ranks (sharded when WORLD > 1)."""
    import torch
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if WORLD > 1:
        import torch.distributed as dist
        V, st, nc, ns = helm_batch_sharded(case, scales, device=torch.device('cuda', torch.cuda.current_device()),
                                           outages=outages, **kw)
        dist.barrier()
    else:
        r = helmpy_b200.helm_batch(case, scales, outages=outages, **kw)
        V, st, nc, ns = r.V, r.status, r.ncoef, r.nswitch
    torch.cuda.synchronize()
    return V, st, nc, ns, time.perf_counter() - t0


def check_vs_oracle(case, V, status, scales, idx, maxc, outages=None, tables=None):
    worst = (0.0, 0.0)
    for b in idx:
        c2 = case
        if outages is not None and outages[b] >= 0:
            from helmpy_b200.core.classes import create_case_data_object_from_arrays
            c2 = create_case_data_object_from_arrays(*tables, 'n1', skip_branch=int(outages[b]))
        Vo, _ = ho.helm_oracle(c2, mismatch=1e-8, scale=float(scales[b]), max_coefficients=maxc)
        assert (status[b] == 2) == (Vo is not None), (b, status[b])
        if Vo is not None:
            e = polar_err(V[b], Vo)
            worst = (max(worst[0], e[0]), max(worst[1], e[1]))
    assert worst[0] <= 1e-8 and worst[1] <= 1e-8, worst
    return {'scenarios_checked': len(idx), 'max_dmag_pu': worst[0], 'max_dang_rad': worst[1]}


def config1():
    out = []
    for name in ('case9', 'case118'):
        case = load_case(name)
        with contextlib.redirect_stdout(io.StringIO()):
            helmpy_b200.helm(case, mismatch=1e-8)
            t0 = time.perf_counter()
            V = helmpy_b200.helm(case, mismatch=1e-8)
            dt = time.perf_counter() - t0
        g = golden_results(name)
        mag, ang = g['classic_mag'], g['classic_ang_deg']
        out.append({'grid': name, 'wall_ms': dt * 1e3, 'max_dmag_pu_vs_reference_golden': float(np.max(np.abs(np.abs(V) - mag))),
                    'max_dang_deg_vs_reference_golden': float(np.max(np.abs(np.degrees(np.angle(V)) - ang)))})
    return {'config': 1, 'what': 'IEEE 118-bus (and case9) HELM-PV2 + Pade single case vs the reference golden files', 'runs': out}


def config2():
    case = load_case('case2869pegase')
    ts = []
    for _ in range(6):
        r = helmpy_b200.helm_batch(case, [1.0], mismatch=1e-8, max_coefficients=40)
        ts.append(r.stats['ms_total'])
    g = golden_results('case2869pegase')
    mag, ang = g['classic_mag'], g['classic_ang_deg']
    V = r.V[0]
    return {'config': 2, 'what': 'case2869pegase HELM-PV2 single case to 1e-8 (CUDA graph)', 'device_ms_median': float(np.median(ts[1:])),
            'coefficients_per_pass': r.list_coef(0), 'switches': int(r.nswitch[0]), 'steps': int(r.stats['steps']),
            'max_dmag_pu_vs_reference_golden': float(np.max(np.abs(np.abs(V) - mag))),
            'max_dang_deg_vs_reference_golden': float(np.max(np.abs(np.degrees(np.angle(V)) - ang)))}


def batch_config(num, what, names, nscen, seed, lo, hi, maxc, ncheck, nr_check=0):
    case = synth.tiled_case(names)
    scales = np.random.default_rng(seed).uniform(lo, hi, nscen)
    kw = dict(mismatch=1e-8, max_coefficients=maxc)
    solve(case, scales, **kw)                                     # warm-up: plan, workspace, page-locked result block
    V, st, nc, ns, dt = solve(case, scales, **kw)
    res = {'config': num, 'what': what, 'grid': 'synthetic: ' + '+'.join(names), 'n_bus': int(case.N), 'scenarios': nscen,
           'n_gpus': WORLD, 'converged': int((st == 2).sum()), 'wall_s': dt, 'cases_per_s': float((st == 2).sum() / dt),
           'mean_passes': float(np.mean((nc > 0).sum(axis=1))), 'mean_switches': float(ns.mean()), 'max_coefficients': maxc}
    if RANK == 0:
        idx = np.random.default_rng(7).choice(nscen, size=ncheck, replace=False)
        res['oracle'] = check_vs_oracle(case, V, st, scales, idx, maxc)
        if nr_check:
            from helmpy_b200.core.nr import nr
            worst = (0.0, 0.0)
            for b in idx[:nr_check]:
                with contextlib.redirect_stdout(io.StringIO()):
                    Vn = nr(case, Mismatch=1e-8, Scale=float(scales[b]))
                assert Vn is not None and st[b] == 2
                e = polar_err(V[b], Vn)
                worst = (max(worst[0], e[0]), max(worst[1], e[1]))
            assert worst[0] <= 1e-6 and worst[1] <= 1e-5, worst
            res['nr_cross_check'] = {'scenarios_checked': nr_check, 'max_dmag_pu': worst[0], 'max_dang_rad': worst[1]}
            # every scenario against the batched GPU Newton-Raphson (csrc/helm_nr.cuh)
            import torch
            helmpy_b200.nr_batch(case, scales[:8], Mismatch=1e-8)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            nrb = helmpy_b200.nr_batch(case, scales, Mismatch=1e-8)
            dtn = time.perf_counter() - t0
            both = nrb.converged & (st == 2)
            dm = float(np.max(np.abs(np.abs(nrb.V[both]) - np.abs(V[both]))))
            da = float(np.max(np.abs(np.angle(nrb.V[both]) - np.angle(V[both]))))
            assert both.sum() == nscen and dm <= 1e-6 and da <= 1e-5, (int(both.sum()), dm, da)
            res['nr_gpu_cross_check'] = {'scenarios_checked': int(both.sum()), 'max_dmag_pu': dm, 'max_dang_rad': da,
                                         'nr_wall_s': dtn, 'nr_cases_per_s': float(nscen / dtn),
                                         'mean_newton_steps': float(np.mean(np.maximum(nrb.iterations - 1, 0).sum(axis=1)))}
    return res


def config3():
    return batch_config(3, 'case2383wp-size grid, 4096 load-scaling scenarios', ['case1354pegase'] + ['case118'] * 8 + ['case9'] * 9,
                        4096, 0, 0.80, 1.05, 60, 8)


def config4():
    return batch_config(4, 'case9241pegase-size grid, PV buses with Q-limit switching, 1024 scenarios, NR cross-check',
                        ['case2869pegase'] * 3 + ['case118'] * 5 + ['case9'] * 5, 1024, 1, 0.85, 1.02, 60, 4, nr_check=2)


def config5():
    names = ['case2869pegase'] * 4 + ['case1354pegase'] + ['case118'] * 7
    tables = synth.tile_tables(names)
    case = synth.tiled_case(names)
    safe = helmpy_b200.non_islanding_branches(case)
    nmax = int(os.environ.get('HELM_N1_MAX', '0'))
    outages = safe if not nmax else safe[:nmax]
    scales = np.ones(len(outages))
    kw = dict(mismatch=1e-8, max_coefficients=60)
    solve(case, scales, outages=outages, **kw)                    # warm-up
    V, st, nc, ns, dt = solve(case, scales, outages=outages, **kw)
    res = {'config': 5, 'what': 'case13659pegase-size N-1 branch-outage sweep (per-scenario numeric refactorization)',
           'grid': 'synthetic: ' + '+'.join(names), 'n_bus': int(case.N), 'branches': int(case.N_branches),
           'non_islanding_outages': int(len(safe)), 'scenarios': int(len(outages)), 'n_gpus': WORLD,
           'converged': int((st == 2).sum()), 'wall_s': dt, 'cases_per_s': float((st == 2).sum() / dt)}
    if RANK == 0:
        idx = np.random.default_rng(7).choice(len(outages), size=2, replace=False)
        res['oracle'] = check_vs_oracle(case, V, st, scales, idx, 60, outages=outages, tables=tables)
        # the outages that did NOT converge: does the reference algorithm (oracle) agree that there is
        # no solution within max_coefficients, or is the GPU path wrong there?
        from helmpy_b200.core.classes import create_case_data_object_from_arrays
        bad = np.nonzero(st != 2)[0]
        verdicts = []
        for b in bad[:int(os.environ.get('HELM_N1_ORACLE_MAX', '24'))]:
            c2 = create_case_data_object_from_arrays(*tables, 'n1', skip_branch=int(outages[b]))
            try:
                Vo, info = ho.helm_oracle(c2, mismatch=1e-8, scale=1.0, max_coefficients=60)
                verdict = 'converged' if Vo is not None else 'diverged'
            except Exception as e:                                # SuperLU: exactly singular matrix
                verdict = 'error:' + type(e).__name__
            verdicts.append({'scenario': int(b), 'branch': int(outages[b]), 'gpu_status': int(st[b]), 'oracle': verdict})
        res['non_converged'] = {'count': int(len(bad)), 'checked': len(verdicts), 'verdicts': verdicts,
                                'oracle_also_without_solution': int(sum(v['oracle'] != 'converged' for v in verdicts))}
        res['non_converged']['same_verdict'] = all(v['oracle'] != 'converged' for v in verdicts)
    return res


def main():
    import torch
    torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
    if WORLD > 1:
        import torch.distributed as dist
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', torch.cuda.current_device()))
    which = [int(a) for a in sys.argv[1:]] or [1, 2, 3, 4, 5]
    fns = {1: config1, 2: config2, 3: config3, 4: config4, 5: config5}
    for k in which:
        if k in (1, 2) and RANK != 0:
            continue
        res = fns[k]()
        if RANK == 0:
            line = json.dumps(res)
            print(line, flush=True)
            os.makedirs(os.path.join(REPO, 'gpurun_out'), exist_ok=True)
            with open(os.path.join(REPO, 'gpurun_out', 'configs.jsonl'), 'a') as f:
                f.write(line + '\n')
    if WORLD > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
