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


def broadcast_factor(core, src=0, group=None):
    """Replicate the factor state of ``core`` (a ``GPCore`` whose hypers and data are already set
    on every rank) from rank ``src``: NCCL broadcast of L, L^-1 (n_pad^2 doubles each) and alpha
    straight out of / into the handle's device buffers, then ``gpc_adopt_factor``."""
    import torch
    import torch.distributed as dist
    pL, pX, pa, n_pad = core.factor_state_dev()
    rank = dist.get_rank(group)
    meta = torch.zeros(2, dtype=torch.float64, device="cuda")
    if rank == src:
        meta[0] = core.logdet_
        meta[1] = 1.0
    torch.cuda.synchronize()
    dist.broadcast(meta, src, group=group)
    for ptr, numel in ((pL, n_pad * n_pad), (pX, n_pad * n_pad), (pa, n_pad)):
        t = _wrap_device_f64(ptr, numel)
        dist.broadcast(t, src, group=group)
    torch.cuda.synchronize()
    if rank != src:
        core.adopt_factor(float(meta[0]))


def _wrap_device_f64(ptr, numel):
    """torch view of a raw device pointer (no copy) through ``__cuda_array_interface__``."""
    import torch

    class _Raw:
        pass

    r = _Raw()
    r.__cuda_array_interface__ = {"shape": (int(numel),), "typestr": "<f8", "data": (int(ptr), False),
                                  "version": 2, "strides": None}
    return torch.as_tensor(r, device="cuda")
