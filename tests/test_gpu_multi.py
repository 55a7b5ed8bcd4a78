"""GPU, two ranks over NCCL (skipped with fewer than two devices): the replicated factor (lower block-triangles of L and
L^-1 broadcast straight into the other rank's handle) predicts bit-identically on both ranks, and the candidate-sharded
information gain -- every rank scores its contiguous share, one 16-byte all-gather picks the winner
(``reduce_best``) -- selects the same node as a single rank scoring all candidates
(replaces the serial scan of ``GraceRIGV3.py:1072-1189``)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu

MF_PARAMS = np.array([3.0, 2.5, 3.5, 3.0, 1.0, 1.5, 2.0, 2.0, 0.5, 1.0, 1.5, 1.5, 0.9, 1.1, 0.08, 0.04, 0.02])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _problem():
    rng = np.random.default_rng(11)
    N, F, C, k = 700, 3, 501, 8
    X4 = np.hstack([rng.uniform([0, 0, 0], [10, 20, 10], (N, 3)), rng.integers(0, F, (N, 1)).astype(float)])
    y = np.sin(X4[:, 0]) + 0.3 * X4[:, 3] + 0.05 * rng.standard_normal(N)
    rows = np.hstack([rng.uniform([0, 0, 0], [10, 20, 10], (C * k, 3)), rng.integers(0, F, (C * k, 1)).astype(float)])
    offs = np.arange(0, (C + 1) * k, k, dtype=np.int64)
    grid4 = np.hstack([rng.uniform([0, 0, 0], [10, 20, 10], (60, 3)), 2.0 * np.ones((60, 1))])
    Xs4 = np.hstack([rng.uniform([0, 0, 0], [10, 20, 10], (3000, 3)), 2.0 * np.ones((3000, 1))])
    return X4, y, rows, offs, grid4, Xs4


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import __graft_entry__ as entry
    entry.setup_path()
    import torch
    import torch.distributed as dist
    import gpcore
    from gpcore import _lib as L
    from gpcore.sharding import broadcast_factor, reduce_best, shard_candidates, shard_range, warm_up
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        warm_up()
        X4, y, rows, offs, grid4, Xs4 = _problem()
        core = gpcore.GPCore(L.KIND_MF_AR1_RBF, 3, rank)
        core.set_hypers(MF_PARAMS, 1e-8)
        core.set_data(X4, y)
        if rank == 0:
            core.factor()
        st = broadcast_factor(core, 0)
        flags = L.INCLUDE_NOISE | L.CLIP_DIAG
        # test points sharded contiguously; rank 0 also predicts everything to compare
        lo, hi = shard_range(len(Xs4), rank, world)
        m, v = core.predict(np.ascontiguousarray(Xs4[lo:hi]), flags)
        parts = [None] * world
        dist.all_gather_object(parts, (lo, hi, m, v))
        # candidates sharded; best node by one 16-byte all-gather
        c_lo, c_hi, loffs, r_lo, r_hi = shard_candidates(offs, rank, world)
        I, _, lbest = core.ig_logdet(grid4, np.ascontiguousarray(rows[r_lo:r_hi]), loffs)
        bv, bi = reduce_best(I[lbest] if lbest >= 0 else 0.0, lbest, c_lo)
        Is, lbs = core.ig_seq(np.ascontiguousarray(rows[r_lo:r_hi]), loffs, float(MF_PARAMS[-1]), pred_fid=0)
        sv, si = reduce_best(Is[lbs] if lbs >= 0 else 0.0, lbs, c_lo)
        if rank == 0:
            mf, vf = core.predict(Xs4, flags)
            If, _, bf = core.ig_logdet(grid4, rows, offs)
            Isf, bsf = core.ig_seq(rows, offs, float(MF_PARAMS[-1]), pred_fid=0)
            same = all(np.array_equal(mf[a:b], pm) and np.array_equal(vf[a:b], pv) for a, b, pm, pv in parts)
            out["res"] = dict(same_prediction=bool(same), best=(bi, int(bf), bv, float(If[bf])),
                              best_seq=(si, int(bsf), sv, float(Isf[bsf])), bytes=st["bytes"],
                              full_bytes=2 * core.padded_n() ** 2 * 8)
        core.close()
    finally:
        dist.destroy_process_group()


def test_two_rank_nccl_factor_broadcast_and_best_candidate(built_lib):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run under gpurun --gpus 2)")
    import torch.multiprocessing as mp
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
        res = dict(out)["res"]
    assert res["same_prediction"]
    bi, bf, bv, vf = res["best"]
    assert bi == bf and bv == vf
    si, bsf, sv, svf = res["best_seq"]
    assert si == bsf and sv == svf
    assert res["bytes"] < 0.7 * res["full_bytes"]           # only the lower block-triangles travelled
