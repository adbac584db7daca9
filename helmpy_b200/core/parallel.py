"""Scenario sharding across the GPUs of one node (one process per GPU).

A scenario (one load scale of one grid) is an independent ``helm`` run
(reference helm.py:896-985 couples nothing across runs), so the batch is split
in contiguous blocks over the ranks and the only exchange is the final gather
of the voltages and per-scenario status — ``torch.distributed`` all_gather:
NCCL over NVLink for CUDA tensors, gloo for the CPU tests.
"""

import numpy as np


def shard_bounds(n_scen, world_size, rank):
    """Contiguous block [lo, hi) of rank; sizes differ by at most one."""
    base, extra = divmod(n_scen, world_size)
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return lo, hi


def gather_blocks(local, n_total, world_size, dist, device=None):
    """All-gather variable-size first-dimension blocks of a numpy array.

    ``local`` is this rank's [n_local, ...] block; returns the [n_total, ...]
    concatenation in rank order on every rank.
    """
    import torch
    local = np.ascontiguousarray(local)
    complex_in = np.iscomplexobj(local)
    arr = local.view(np.float64).reshape(local.shape + (2,)) if complex_in else local
    base, extra = divmod(n_total, world_size)
    nmax = base + (1 if extra else 0)
    pad_shape = (nmax,) + arr.shape[1:]
    buf = np.zeros(pad_shape, dtype=arr.dtype)
    buf[:arr.shape[0]] = arr
    t = torch.from_numpy(buf)
    if device is not None:
        t = t.to(device)
    out = [torch.empty_like(t) for _ in range(world_size)]
    dist.all_gather(out, t)
    parts = []
    for r in range(world_size):
        lo, hi = shard_bounds(n_total, world_size, r)
        parts.append(out[r][:hi - lo].cpu().numpy())
    full = np.concatenate(parts, axis=0)
    if complex_in:
        full = np.ascontiguousarray(full).view(np.complex128).reshape(full.shape[:-1])
    return full


def assemble_gathered(Vall, Iall, n, world, nmax, max_passes):
    """Full (V, status, ncoef, nswitch) from the gathered buffers: Vall complex [world*nmax, N],
    Iall int32 [world, 2 + max_passes, nmax] (rows: status, nswitch, ncoef of every pass); rank r
    holds scenarios shard_bounds(n, world, r) in its first entries, the rest is padding."""
    if n == world * nmax:                                    # equal blocks: the gathered buffer IS the result
        status = np.ascontiguousarray(Iall[:, 0, :]).reshape(n)
        nswitch = np.ascontiguousarray(Iall[:, 1, :]).reshape(n)
        ncoef = np.ascontiguousarray(Iall[:, 2:, :].transpose(0, 2, 1)).reshape(n, max_passes)
        return Vall, status, ncoef, nswitch
    Vp, sp, cp, wp = [], [], [], []
    for r in range(world):
        rlo, rhi = shard_bounds(n, world, r)
        k = rhi - rlo
        Vp.append(Vall[r * nmax: r * nmax + k])
        sp.append(Iall[r, 0, :k])
        wp.append(Iall[r, 1, :k])
        cp.append(Iall[r, 2:, :k].T)
    return np.concatenate(Vp), np.concatenate(sp), np.ascontiguousarray(np.concatenate(cp)), np.concatenate(wp)


def _sharded_device(case, scales, outages, device, root, kw, device_solver=None, n_bus=None, max_passes=16):
    """Device-resident path: each rank solves its block straight into device buffers
    (helm_solve_batch_device), then ONE all_gather_into_tensor for the voltages and one for the
    per-scenario integers (NCCL over NVLink / NVSwitch) and one device-to-host copy of the result
    into page-locked memory.  No numpy bounce between solve and gather.

    ``device_solver(d_scale, d_V, d_status, d_ncoef, d_nswitch, outages_block)`` replaces the CUDA
    call in the CPU tests (gloo, CPU tensors); the buffer handling is the same code.
    """
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    n = len(scales)
    lo, hi = shard_bounds(n, world, rank)
    nloc = hi - lo
    nmax = -(-n // world)
    is_cuda = getattr(device, 'type', str(device)) == 'cuda'
    plan = None
    if device_solver is None:
        from . import _lib
        from .helm import get_plan
        plan = get_plan(case)
        n_bus = plan.N
        max_passes = _lib.HELM_MAX_PASSES
        dsb_method = kw.get('DSB_model_method')
        if kw.get('DSB_model') and dsb_method is None:
            dsb_method = 2
        prm = _lib.Plan.make_params(kw.get('mismatch', 1e-4), kw.get('max_coefficients', 100),
                                    kw.get('enforce_Q_limits', True), kw.get('group', 0), kw.get('use_graph', True),
                                    kw.get('poll_every', 0), kw.get('max_resident', 0), kw.get('pv_bus_model', 2),
                                    dsb_method or 0, kw.get('DSB_model', False))
    N, MAXP = n_bus, max_passes
    holder = plan if plan is not None else _sharded_device
    key = (nmax, world, str(device), N)
    bufs = getattr(holder, '_shard_bufs', None)
    if bufs is None or bufs['key'] != key:
        bufs = {'key': key,
                'scale': torch.empty(nmax, dtype=torch.float64, device=device),
                'V': torch.zeros(nmax, N, dtype=torch.complex128, device=device),
                'ints': torch.zeros(2 + MAXP, nmax, dtype=torch.int32, device=device),   # rows: status, nswitch, ncoef of each pass
                'ncoef': torch.zeros(nmax, MAXP, dtype=torch.int32, device=device),
                'gV': torch.empty(world * nmax, N, 2, dtype=torch.float64, device=device),
                'gI': torch.empty(world * (2 + MAXP), nmax, dtype=torch.int32, device=device),
                'ws': None,
                # the solve replays a captured CUDA graph: it needs a stream of its own (capture is not
                # allowed on the legacy default stream)
                'stream': torch.cuda.Stream(device) if is_cuda else None}
        holder._shard_bufs = bufs
    ints = bufs['ints']
    if nloc:
        bufs['scale'][:nloc].copy_(torch.from_numpy(np.ascontiguousarray(scales[lo:hi])))
        ob = None if outages is None else np.ascontiguousarray(outages[lo:hi], dtype=np.int32)
        if device_solver is not None:
            device_solver(bufs['scale'][:nloc], bufs['V'][:nloc], ints[0, :nloc], bufs['ncoef'][:nloc], ints[1, :nloc], ob)
        else:
            import ctypes as C
            need = plan.workspace_bytes(nloc, prm)
            if bufs['ws'] is None or bufs['ws'].numel() < need:
                bufs['ws'] = None
                bufs['ws'] = torch.empty(need, dtype=torch.uint8, device=device)
            if ob is not None:
                _lib.check(_lib.load().helm_plan_set_outages(plan.handle, nloc, ob.ctypes.data_as(C.c_void_p)),
                           "helm_plan_set_outages")
            st = bufs['stream']
            st.wait_stream(torch.cuda.current_stream(device))
            with torch.cuda.stream(st):
                plan.solve_device(bufs['scale'][:nloc], prm, bufs['ws'], bufs['V'][:nloc], ints[0, :nloc],
                                  bufs['ncoef'][:nloc], ints[1, :nloc], st.cuda_stream)
            torch.cuda.current_stream(device).wait_stream(st)
        ints[2:, :nloc].copy_(bufs['ncoef'][:nloc].t())
    dist.all_gather_into_tensor(bufs['gV'], torch.view_as_real(bufs['V']))
    dist.all_gather_into_tensor(bufs['gI'], ints)
    if root is not None and rank != root:
        if is_cuda:
            torch.cuda.current_stream(device).synchronize()
        return None, None, None, None
    hV = torch.empty(bufs['gV'].shape, dtype=torch.float64, pin_memory=is_cuda)
    hI = torch.empty(bufs['gI'].shape, dtype=torch.int32, pin_memory=is_cuda)
    hV.copy_(bufs['gV'], non_blocking=is_cuda)
    hI.copy_(bufs['gI'], non_blocking=is_cuda)
    if is_cuda:
        torch.cuda.current_stream(device).synchronize()
    Vall = hV.numpy().view(np.complex128).reshape(world * nmax, N)
    Iall = hI.numpy().reshape(world, 2 + MAXP, nmax)
    return assemble_gathered(Vall, Iall, n, world, nmax, MAXP)


def helm_batch_sharded(case, scales, solve_fn=None, device=None, root=None, device_solver=None, **kw):
    """Solve ``scales`` across all ranks of the default process group.

    Every rank passes the same ``scales`` (and ``outages``, if any); each solves its contiguous
    block and all ranks return the full (V, status, ncoef, nswitch) — or only rank ``root`` does
    (the others get four ``None``) when ``root`` is given.  On CUDA devices the block is solved into
    device buffers and gathered with one NCCL all-gather (``_sharded_device``).  With ``solve_fn``
    (``solve_fn(case, scales_block, **kw)`` returning an object with ``V``, ``status``, ``ncoef``,
    ``nswitch``; used by the gloo tests on CPU) the blocks are host arrays and go through
    ``gather_blocks``.
    """
    import torch.distributed as dist
    from .helm import helm_batch
    world, rank = dist.get_world_size(), dist.get_rank()
    outages = kw.pop('outages', None)
    if scales is None:
        scales = np.ones(len(outages))
    scales = np.asarray(scales, dtype=np.float64)
    if outages is not None:
        outages = np.asarray(outages, dtype=np.int32)
    if device_solver is not None:                # CPU tests of the device-resident path
        return _sharded_device(case, scales, outages, device, root, kw, device_solver=device_solver,
                               n_bus=kw.pop('n_bus'), max_passes=kw.pop('max_passes', 16))
    if solve_fn is None and device is not None and getattr(device, 'type', str(device)) == 'cuda':
        return _sharded_device(case, scales, outages, device, root, kw)
    lo, hi = shard_bounds(len(scales), world, rank)
    fn = solve_fn or helm_batch
    if outages is not None:                      # N-1 scenarios are sharded with their scales
        kw['outages'] = outages[lo:hi]
    res = fn(case, scales[lo:hi], **kw)
    n = len(scales)
    V = gather_blocks(res.V, n, world, dist, device)
    status = gather_blocks(res.status, n, world, dist, device)
    ncoef = gather_blocks(res.ncoef, n, world, dist, device)
    nswitch = gather_blocks(res.nswitch, n, world, dist, device)
    if root is not None and rank != root:
        return None, None, None, None
    return V, status, ncoef, nswitch
