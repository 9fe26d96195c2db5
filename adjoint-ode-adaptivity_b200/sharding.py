"""Multi-GPU plumbing: the batch of trajectories shards over ranks with no data-path
collective; the only exchange is the all-reduce of the per-rank indicator
partials that feeds the batch-mean refinement rule of the reference
(python/Main_variable_params.py:340-341: jnp.mean(err, axis=0) -> argmax)."""
from __future__ import annotations

import numpy as np


def shard_range(B: int, rank: int, world: int):
    """Contiguous slice [lo, hi) of a global batch of B trajectories owned by `rank`."""
    base, rem = divmod(B, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_indicators(sums, group=None, ordered=True):
    """Combine per-rank partials sums[K+4] = [sum_b|eta[b,k]| (K), sum|eta|, sum eta^2,
    max|eta|, sum J] across ranks.  ordered=True all-gathers the partials and adds them in
    rank order on every rank: the same bits on every rank whatever the reduction topology
    (SURVEY hard part 6); ordered=False is a plain all-reduce (sum + max).  One partial per rank
    still makes the batch sum depend on the number of ranks at rounding level -- use
    `allreduce_indicator_blocks` for results that are bit-identical across 1/2/4/8 GPUs."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return sums
    world = dist.get_world_size(group)
    K = sums.numel() - 4
    if ordered:
        parts = [torch.empty_like(sums) for _ in range(world)]
        dist.all_gather(parts, sums, group=group)
        out = parts[0].clone()
        for p in parts[1:]:
            out[:K + 2] += p[:K + 2]
            out[K + 2] = torch.maximum(out[K + 2], p[K + 2])
            out[K + 3] += p[K + 3]
        return out
    out = sums.clone()
    mx = out[K + 2:K + 3].clone()
    dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
    out[K + 2] = mx[0]
    return out


REDUCE_BLOCK = 4096   # trajectories per reduction block of the count-independent batch sums


def combine_blocks(parts):
    """Sum the rows of parts[nblk, K+4] in row order (entry K+2: a maximum).  CUDA tensors go through
    the C-ABI (`dgadj_allreduce_indicator_blocks` without a communicator); CPU tensors (the gloo
    tests of the host logic) are added row by row here -- same order, same bits."""
    import torch
    K = parts.shape[1] - 4
    if parts.is_cuda:
        from . import _lib
        import ctypes as C
        lib = _lib.load()
        h = _any_handle(parts.device.index)
        out = torch.empty(K + 4, dtype=torch.float64, device=parts.device)
        parts = parts.contiguous()
        rc = lib.dgadj_allreduce_indicator_blocks(h, C.c_void_p(0), K, parts.shape[0], C.c_void_p(parts.data_ptr()),
                                                  C.c_void_p(out.data_ptr()),
                                                  C.c_void_p(torch.cuda.current_stream(parts.device).cuda_stream))
        if rc != _lib.OK:
            raise _lib.DgadjError(rc, lib.dgadj_last_error(h).decode())
        return out
    out = parts[0].clone()
    for r in range(1, parts.shape[0]):
        mx = torch.maximum(out[K + 2], parts[r, K + 2])
        out += parts[r]
        out[K + 2] = mx
    return out


_HANDLES = {}


def _any_handle(device):
    """A minimal handle on `device` for calls that need none of a solver's state."""
    import ctypes as C
    from . import _lib
    if device not in _HANDLES:
        lib = _lib.load()
        cfg = _lib.Config(device=device, N=1, K=1, bc=1, inflow=0, functional=0, scheme=0, reserved=0, alpha=0.0)
        h = C.c_void_p(0)
        rc = lib.dgadj_create(C.byref(cfg), C.byref(h))
        if rc != _lib.OK:
            raise _lib.DgadjError(rc, "dgadj_create failed (an sm_100 device is required; there is no CPU path)")
        _HANDLES[device] = h
    return _HANDLES[device]


def allreduce_indicator_blocks(parts, group=None):
    """Count-independent batch sums: `parts[nblk_local, K+4]` are this rank's block partials
    (`AdvecDG1D.reduce_indicator_blocks`: fixed blocks of REDUCE_BLOCK trajectories); the rows of all
    ranks are all-gathered in rank order (= global block order for `shard_range` shards) and added in
    that order on every rank.  When every shard is a whole number of blocks the result has the same bits
    on 1, 2, 4 or 8 GPUs, so near-ties of the batch-mean argmax refine the same element everywhere."""
    import torch
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        world = dist.get_world_size(group)
        rows = [torch.empty_like(parts) for _ in range(world)]
        dist.all_gather(rows, parts.contiguous(), group=group)
        parts = torch.cat(rows, dim=0)
    return combine_blocks(parts)


def gather_indicators(eta, group=None):
    """All-gather the per-rank indicator slices eta[B_r, K] into the global [B, K] array on every
    rank (rank order = batch order of `shard_range`); used when per-trajectory rankings are wanted
    in one place.  Slices may have different lengths."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return eta
    world = dist.get_world_size(group)
    n = torch.tensor([eta.shape[0]], dtype=torch.int64, device=eta.device)
    counts = [torch.empty_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c[0]) for c in counts]
    nmax = max(counts)
    pad = eta if eta.shape[0] == nmax else torch.cat([eta, eta.new_zeros((nmax - eta.shape[0],) + eta.shape[1:])])
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad.contiguous(), group=group)
    return torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)


def batch_mean_refine(sums, B_global: int):
    """Shared-mesh refinement decision from the reduced partials: mean indicator per element
    and the element to refine (argmax, lowest index on ties -- np.argmax semantics,
    python/Main_finite_difference.py:337; python/Main_variable_params.py:340-341)."""
    s = sums.detach().cpu().numpy() if hasattr(sums, "detach") else np.asarray(sums)
    K = s.size - 4
    mean_ind = s[:K] / float(B_global)
    return mean_ind, int(np.argmax(mean_ind))


def allreduce_indicators_comm(solver, nccl_comm, sums):
    """The same ordered combination through the C-ABI (`dgadj_allreduce_indicators`) on a raw
    NCCL communicator (an `ncclComm_t` as an integer / ctypes pointer) -- what a C, C++ or MPI
    host of the library calls; `sums` (float64 CUDA tensor [K+4]) is updated in place."""
    import ctypes as C
    import torch
    from . import _lib
    K = sums.numel() - 4
    rc = solver.lib.dgadj_allreduce_indicators(solver._h, C.c_void_p(int(nccl_comm)), K, C.c_void_p(sums.data_ptr()),
                                               C.c_void_p(torch.cuda.current_stream(sums.device).cuda_stream))
    if rc != _lib.OK:
        raise _lib.DgadjError(rc, solver.lib.dgadj_last_error(solver._h).decode())
    return sums
