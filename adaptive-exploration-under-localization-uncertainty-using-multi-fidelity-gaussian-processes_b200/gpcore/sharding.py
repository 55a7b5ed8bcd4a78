"""One process per GPU: the factor is replicated, test points and candidates are sharded
(SURVEY.md section 8e).  ``torch.distributed`` is plumbing only -- NCCL on the GPUs (broadcast
of L / L^-1 / alpha from the factoring rank, all-gather of 16 bytes per rank for the best
candidate), gloo in the CPU tests of this host logic.  There is no collective inside the
compute: each rank runs the same kernels on its own contiguous slice.
"""
import numpy as np


def shard_range(total, rank, world):
    """Contiguous slice [lo, hi) of ``total`` items owned by ``rank``; sizes differ by at most 1
    and earlier ranks take the extra items."""
    total, rank, world = int(total), int(rank), int(world)
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_candidates(offsets, rank, world):
    """Slice a ragged candidate set by candidate count.  Returns (c_lo, c_hi, local_offsets,
    row_lo, row_hi)."""
    offsets = np.asarray(offsets, dtype=np.int64)
    c_lo, c_hi = shard_range(offsets.size - 1, rank, world)
    return c_lo, c_hi, offsets[c_lo:c_hi + 1] - offsets[c_lo], int(offsets[c_lo]), int(offsets[c_hi])


def reduce_best(local_best_value, local_best_index, index_offset, group=None):
    """Gather (max I, global argmax) over ranks: one all-gather of two doubles per rank.
    NaN / empty shards pass ``local_best_index < 0``.  Ties resolve to the lowest global index,
    matching a single-process argmax over the concatenated scores."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else "cpu"
    gidx = float(local_best_index + index_offset) if local_best_index >= 0 else -1.0
    mine = torch.tensor([float(local_best_value) if local_best_index >= 0 else float("-inf"), gidx],
                        dtype=torch.float64, device=dev)
    allv = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allv, mine, group=group)
    best_v, best_i = float("-inf"), -1
    for t in allv:
        v, i = float(t[0]), int(t[1])
        if i >= 0 and (best_i < 0 or v > best_v or (v == best_v and i < best_i)):
            best_v, best_i = v, i
    return best_v, best_i


def trapezoids(n_pad, groups=8, block=128):
    """Cover the lower block-triangle of an n_pad x n_pad matrix with ``groups`` row bands [r0, r1) x [0, r1): what
    ``broadcast_factor`` sends instead of the full squares (the upper triangles of L and L^-1 are zeros).  Bands are
    multiples of ``block`` rows; returns [(r0, r1), ...] and the fraction of n_pad^2 they hold (0.5 + 0.5 / groups)."""
    nb = int(n_pad) // block
    groups = max(1, min(int(groups), nb))
    cuts = [round(g * nb / groups) * block for g in range(groups + 1)]
    bands = [(cuts[g], cuts[g + 1]) for g in range(groups) if cuts[g + 1] > cuts[g]]
    frac = sum((r1 - r0) * r1 for r0, r1 in bands) / float(n_pad * n_pad)
    return bands, frac


_STAGING = {}


def _staging(numel, like):
    """One reusable contiguous staging tensor per (device, dtype), grown on demand: the bands of a factor broadcast go
    through it one after the other, so a replication costs one allocation ever instead of one per band per call."""
    key = (str(like.device), like.dtype)
    buf = _STAGING.get(key)
    if buf is None or buf.numel() < numel:
        import torch
        buf = _STAGING[key] = torch.empty(int(numel), dtype=like.dtype, device=like.device)
    return buf


def broadcast_lower(t2d, src=0, group=None, groups=8):
    """Broadcast the lower block-triangle of the square tensor ``t2d`` (any backend) band by band: the band is packed
    into a contiguous staging buffer, broadcast, and written back on the receivers.  Returns the bytes broadcast."""
    import torch.distributed as dist
    rank = dist.get_rank(group)
    n = t2d.shape[0]
    bands, _ = trapezoids(n, groups)
    stage = _staging(max((r1 - r0) * r1 for r0, r1 in bands), t2d)
    sent = 0
    for r0, r1 in bands:
        view = t2d[r0:r1, :r1]
        buf = stage[:(r1 - r0) * r1].view(r1 - r0, r1)
        if rank == src:
            buf.copy_(view)
        dist.broadcast(buf, src, group=group)
        if rank != src:
            view.copy_(buf)
        sent += buf.numel() * buf.element_size()
    return sent


def warm_up(group=None):
    """Create the communicator before anything is timed: the first NCCL collective of a process group pays for the
    communicator set-up (~1 s at 8 ranks), which is not part of any factor broadcast."""
    import torch
    import torch.distributed as dist
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.ones(1, device=dev)
    dist.all_reduce(t, group=group)
    dist.broadcast(t, 0, group=group)
    if dev != "cpu":
        # large messages take another protocol / more channels than a 4-byte one: connect those too
        big = torch.empty(8 << 20, dtype=torch.float64, device=dev)
        dist.broadcast(big, 0, group=group)
        dist.broadcast(big, 0, group=group)
        torch.cuda.synchronize()


def broadcast_factor(core, src=0, group=None, triangular=True, groups=8):
    """Replicate the factor state of ``core`` (a ``GPCore`` whose hypers and data are already set
    on every rank) from rank ``src``: NCCL broadcast of L, L^-1 (their lower block-triangles when ``triangular``,
    0.5 + 0.5 / groups of the n_pad^2 doubles each) and alpha straight out of / into the handle's device buffers,
    then ``gpc_adopt_factor``.  Returns {"bytes", "ms", "gbps"} (time between device synchronisations on this rank)."""
    import time
    import torch
    import torch.distributed as dist
    pL, pX, pa, n_pad = core.factor_state_dev()
    rank = dist.get_rank(group)
    meta = torch.zeros(2, dtype=torch.float64, device="cuda")
    if rank == src:
        meta[0] = core.logdet_
        meta[1] = 1.0
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    dist.broadcast(meta, src, group=group)
    sent = 16
    for ptr in (pL, pX):
        t = _wrap_device_f64(ptr, n_pad * n_pad)
        if triangular:
            if rank != src:
                t.zero_()                       # the upper triangle is never sent
            sent += broadcast_lower(t.view(n_pad, n_pad), src, group, groups)
        else:
            dist.broadcast(t, src, group=group)
            sent += n_pad * n_pad * 8
    dist.broadcast(_wrap_device_f64(pa, n_pad), src, group=group)
    sent += n_pad * 8
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if rank != src:
        core.adopt_factor(float(meta[0]))
    return {"bytes": int(sent), "ms": 1e3 * dt, "gbps": sent / dt / 1e9}


REDUNDANT_FACTOR_MAX_N = 4096


def replicate_factor(core, src=0, group=None, redundant_max_n=REDUNDANT_FACTOR_MAX_N):
    """Every rank ends up with the factor.  Up to ``redundant_max_n`` training points each rank factors its own copy
    (2 ms at N = 2048, 5 ms at 4096: cheaper than moving 2 x 8 N^2 bytes and free of any collective -- SURVEY 8e);
    beyond that rank ``src`` factors and the lower block-triangles are broadcast.
    Returns {"mode", "factor_ms", "bytes", "ms", "gbps"}."""
    import time
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if world == 1 or core.N <= redundant_max_n:
        core.factor()
        return {"mode": "redundant" if world > 1 else "single", "factor_ms": 1e3 * (time.perf_counter() - t0),
                "bytes": 0, "ms": 0.0, "gbps": None}
    if dist.get_rank(group) == src:
        core.factor()
    t_factor = time.perf_counter() - t0
    st = broadcast_factor(core, src, group)
    st.update(mode="broadcast_lower_triangles", factor_ms=1e3 * t_factor)
    return st


def _wrap_device_f64(ptr, numel):
    """torch view of a raw device pointer (no copy) through ``__cuda_array_interface__``."""
    import torch

    class _Raw:
        pass

    r = _Raw()
    r.__cuda_array_interface__ = {"shape": (int(numel),), "typestr": "<f8", "data": (int(ptr), False),
                                  "version": 2, "strides": None}
    return torch.as_tensor(r, device="cuda")
