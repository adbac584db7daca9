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


def helm_batch_sharded(case, scales, solve_fn=None, device=None, **kw):
    """Solve ``scales`` across all ranks of the default process group.

    Every rank passes the same ``scales`` (and ``outages``, if any); each solves its block with
    ``helm_batch`` (or ``solve_fn(case, scales_block, **kw)`` returning an object
    with ``V``, ``status``, ``ncoef``, ``nswitch``) and all ranks return the
    full (V, status, ncoef, nswitch).
    """
    import torch.distributed as dist
    from .helm import helm_batch
    world, rank = dist.get_world_size(), dist.get_rank()
    outages = kw.pop('outages', None)
    if scales is None:
        scales = np.ones(len(outages))
    scales = np.asarray(scales, dtype=np.float64)
    lo, hi = shard_bounds(len(scales), world, rank)
    fn = solve_fn or helm_batch
    if outages is not None:                      # N-1 scenarios are sharded with their scales
        kw['outages'] = np.asarray(outages, dtype=np.int32)[lo:hi]
    res = fn(case, scales[lo:hi], **kw)
    n = len(scales)
    V = gather_blocks(res.V, n, world, dist, device)
    status = gather_blocks(res.status, n, world, dist, device)
    ncoef = gather_blocks(res.ncoef, n, world, dist, device)
    nswitch = gather_blocks(res.nswitch, n, world, dist, device)
    return V, status, ncoef, nswitch
