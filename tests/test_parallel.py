"""CPU, world_size 2 over gloo: the scenario-sharding host path (block split +
final gather) with a stub solver in place of the CUDA call."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from helmpy_b200.core import parallel


def test_shard_bounds_cover_everything():
    for n in (1, 7, 16, 4096, 4097):
        for w in (1, 2, 3, 8):
            got = []
            for r in range(w):
                lo, hi = parallel.shard_bounds(n, w, r)
                got += list(range(lo, hi))
            assert got == list(range(n))


class _Res:
    pass


def _stub_solver(case, scales, outages=None, **kw):
    r = _Res()
    n = len(scales)
    r.V = (np.outer(scales, np.arange(1, 6)) + 1j * scales[:, None]).astype(np.complex128)
    if outages is not None:                    # N-1 scenarios: the removed branch must travel with its scale
        assert len(outages) == n
        r.V = r.V + 1000.0 * np.asarray(outages, dtype=np.float64)[:, None]
    r.status = np.full(n, 2, dtype=np.int32)
    r.ncoef = np.tile(np.arange(16, dtype=np.int32), (n, 1)) * (1 + np.arange(n, dtype=np.int32))[:, None]
    r.nswitch = (scales * 100).astype(np.int32)
    return r


def _stub_device_solver(d_scale, d_V, d_status, d_ncoef, d_nswitch, outages):
    """Same results as _stub_solver, written into the caller's (CPU) torch buffers the way
    helm_solve_batch_device writes device memory."""
    import torch
    sc = d_scale.numpy().copy()
    r = _stub_solver(None, sc, outages=outages)
    d_V.copy_(torch.from_numpy(r.V))
    d_status.copy_(torch.from_numpy(r.status))
    d_ncoef.copy_(torch.from_numpy(r.ncoef))
    d_nswitch.copy_(torch.from_numpy(r.nswitch))


def _worker(rank, world, port, n, q):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    scales = np.linspace(0.8, 1.05, n)
    V, status, ncoef, nswitch = parallel.helm_batch_sharded(None, scales, solve_fn=_stub_solver)
    ref = _stub_solver(None, scales)
    ok = (np.array_equal(V, ref.V) and np.array_equal(status, ref.status)
          and np.array_equal(nswitch, ref.nswitch) and V.dtype == np.complex128 and V.shape == (n, 5))
    # ncoef of the stub depends on the local index; rebuild the expected concatenation
    exp = []
    for r in range(world):
        lo, hi = parallel.shard_bounds(n, world, r)
        exp.append(_stub_solver(None, scales[lo:hi]).ncoef)
    ok = ok and np.array_equal(ncoef, np.concatenate(exp))
    # N-1 sweep: outages are sharded with the scales (scales=None means all ones)
    outages = np.arange(n, dtype=np.int32)[::-1].copy()
    V2, st2, _, _ = parallel.helm_batch_sharded(None, None, solve_fn=_stub_solver, outages=outages)
    ok = ok and np.array_equal(V2, _stub_solver(None, np.ones(n), outages=outages).V)
    # the device-resident path (one all_gather_into_tensor of padded blocks) with CPU tensors
    import torch
    for root in (None, 0):
        out = parallel.helm_batch_sharded(None, scales, device=torch.device('cpu'), device_solver=_stub_device_solver,
                                          root=root, n_bus=5, outages=outages)
        if root is not None and rank != root:
            ok = ok and all(o is None for o in out)
            continue
        V3, st3, nc3, ns3 = out
        ref3 = _stub_solver(None, scales, outages=outages)
        ok = ok and np.array_equal(V3, ref3.V) and np.array_equal(st3, ref3.status) and np.array_equal(ns3, ref3.nswitch)
        ok = ok and np.array_equal(nc3, np.concatenate(exp)) and V3.shape == (n, 5) and nc3.shape == (n, 16)
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


@pytest.mark.parametrize('n', [7, 16])
def test_sharded_gather_gloo_world2(n):
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
    results = sorted(q.get(timeout=10) for _ in range(2))
    assert results == [(0, True), (1, True)]
