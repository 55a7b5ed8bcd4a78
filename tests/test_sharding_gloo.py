"""CPU, world_size 2 over gloo: the N > 1 host logic -- contiguous sharding of test points and
candidates, and the 16-byte-per-rank gather that selects the best candidate."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, scores, out):
    sys.path.insert(0, ROOT)
    import __graft_entry__ as entry
    entry.setup_path()
    import torch.distributed as dist
    from gpcore.sharding import reduce_best, shard_range
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        for case, sc in enumerate(scores):
            lo, hi = shard_range(len(sc), rank, world)
            loc = np.asarray(sc[lo:hi], dtype=float)
            ok = ~np.isnan(loc)
            if ok.any():
                li = int(np.argmax(np.where(ok, loc, -np.inf)))
                v, i = reduce_best(loc[li], li, lo)
            else:
                v, i = reduce_best(0.0, -1, lo)
            out[(case, rank)] = (v, i)
        # lower block-triangle broadcast (what replicates L and L^-1 across ranks), on CPU tensors over gloo
        import torch
        from gpcore.sharding import broadcast_lower, trapezoids, warm_up
        warm_up()
        n = 1024
        full = torch.tril(torch.arange(n * n, dtype=torch.float64).reshape(n, n) + 1.0)
        t = full.clone() if rank == 0 else torch.zeros(n, n, dtype=torch.float64)
        sent = broadcast_lower(t, 0, None, groups=4)
        bands, frac = trapezoids(n, 4)
        out[("tri", rank)] = (bool(torch.equal(t, full)), sent, frac)
    finally:
        dist.destroy_process_group()


def test_best_candidate_gather_world2():
    rng = np.random.default_rng(0)
    scores = [rng.standard_normal(101).tolist(),
              [1.0, 5.0, 5.0, 2.0],                      # tie across ranks -> lowest global index
              [float("nan"), float("nan"), 3.0],         # a rank whose shard is all NaN
              [2.0]]                                      # fewer candidates than ranks
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(2, _free_port(), scores, out), nprocs=2, join=True)
        out = dict(out)
    for rank in range(2):
        same, sent, frac = out[("tri", rank)]
        assert same and sent == int(frac * 1024 * 1024) * 8 and 0.5 < frac <= 0.5 + 0.5 / 4 + 1e-12
    for case, sc in enumerate(scores):
        a = np.asarray(sc, float)
        want = int(np.nanargmax(a))
        for rank in range(2):
            v, i = out[(case, rank)]
            assert i == want and v == a[want], (case, rank, v, i)


def test_trapezoid_cover():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as entry
    entry.setup_path()
    from gpcore.sharding import trapezoids
    for n_pad, groups in ((128, 8), (2048, 8), (8192, 8), (16384, 16), (640, 3)):
        bands, frac = trapezoids(n_pad, groups)
        assert bands[0][0] == 0 and bands[-1][1] == n_pad
        assert all(a[1] == b[0] for a, b in zip(bands, bands[1:])) and all(r0 % 128 == 0 for r0, _ in bands)
        covered = np.zeros((n_pad // 128, n_pad // 128), bool)
        for r0, r1 in bands:
            covered[r0 // 128:r1 // 128, :r1 // 128] = True
        assert np.all(covered[np.tril_indices(n_pad // 128)])           # every lower block is sent
        assert frac <= 0.5 + 0.5 / min(groups, n_pad // 128) + 0.5 / (n_pad // 128) + 1e-12
