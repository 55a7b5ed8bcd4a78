"""BASELINE configs[4]: scaling sweep over the training-set size N and the test-set size M, single- and
multi-fidelity, one process per GPU (run under torchrun for more than one GPU; test points are sharded, the
factor is replicated).  Prints one JSON line per cell: posterior mean + variance points/s (device-resident
inputs, CUDA events on the library's stream), the factorisation time, and the FP64-equivalent rate of the
dominant contraction (N^2 flop per point) against the measured DGEMM peak.

    python examples/sweep.py                      # 1 GPU, default cells
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 examples/sweep.py

M is capped per cell by --max-points (throughput does not depend on M beyond a few 16384-row chunks); the cap
is stated in the line."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

entry.setup_path()
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-train", type=int, nargs="*", default=[1024, 2048, 4096, 8192, 16384])
    ap.add_argument("--log2-m", type=int, nargs="*", default=[20, 24])
    ap.add_argument("--max-points", type=float, default=3.0e13,
                    help="cap on M * N^2 per cell per GPU (FP64-equivalent flop), default = a few seconds")
    ap.add_argument("--mode", default="int8", choices=["int8", "fp64"])
    args = ap.parse_args()
    import torch
    import gpcore
    from gpcore import _lib as L
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    dist = None
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ.pop("NCCL_DEBUG", None)     # keep stdout to the JSON lines (NCCL prints its version banner there)
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    gpcore.build()
    peak = 35.46
    try:
        peak = float(json.load(open(bench.DGEMM_PEAK_FILE))["dgemm_8192_tflops"])
    except Exception:
        pass
    mode = L.MODE_INT8 if args.mode == "int8" else L.MODE_FP64
    for kind, F, params in (("sf_rbf", 1, np.array([4.0, 2.0, 3.0, 2.5, 0.05])), ("mf_ar1", 2, bench.MF2_PARAMS)):
        for N in args.n_train:
            X4, y = bench.make_train(N, F)
            core = gpcore.GPCore(L.KIND_SF_RBF if F == 1 else L.KIND_MF_AR1_RBF, F, local)
            core.set_hypers(params, 1e-8)
            core.set_data(X4, y)
            core.set_mode(mode)
            core.factor()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            core.factor()
            t_factor = time.perf_counter() - t0
            stream = torch.cuda.ExternalStream(core.stream())
            for lg in args.log2_m:
                M_total = 1 << lg
                M = M_total // world
                cap = int(args.max_points / (float(N) * N))
                Mrun = max(16384, min(M, cap) // 16384 * 16384)
                side = int(round(Mrun ** (1 / 3))) + 1
                Xs = bench.make_grid(side, F - 1)[:Mrun]
                dXs = torch.from_numpy(Xs).cuda()
                dmean = torch.empty(Mrun, dtype=torch.float64, device="cuda")
                dvar = torch.empty(Mrun, dtype=torch.float64, device="cuda")
                flags = L.INCLUDE_NOISE | L.CLIP_DIAG
                core.predict_dev(dXs.data_ptr(), Mrun, dmean.data_ptr(), dvar.data_ptr(), flags)
                torch.cuda.synchronize()
                if dist is not None:
                    dist.barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                core.predict_dev(dXs.data_ptr(), Mrun, dmean.data_ptr(), dvar.data_ptr(), flags)
                e1.record(stream)
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1)
                if dist is not None:
                    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    ms = float(t[0])
                pts = world * Mrun / (ms * 1e-3)
                if rank == 0:
                    print(json.dumps({"kind": kind, "n_train": N, "m_test_total": M_total, "n_gpus": world,
                                      "m_timed_per_gpu": Mrun, "capped": Mrun < M, "pts_per_s": pts,
                                      "ms": ms, "factor_ms": 1e3 * t_factor, "mode": args.mode,
                                      "fp64_equiv_tflops_per_gpu": pts / world * float(N) * N / 1e12,
                                      "over_dgemm_peak": pts / world * float(N) * N / 1e12 / peak,
                                      "finite": bool(torch.isfinite(dvar).all().item())}), flush=True)
                del dXs, dmean, dvar
            core.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
