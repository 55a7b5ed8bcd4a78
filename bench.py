#!/usr/bin/env python
"""bench.py -- GP posterior mean+var throughput (BASELINE.json metric) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload ...]

Workload at the default settings = BASELINE.json configs[1]: synthetic 2-fidelity AR1 GP,
N = 2048 training points, 100 x 100 x 100 test grid (1 M points) at the top fidelity, fixed
hyper-parameters.  One *step* = one pass of the posterior (mean + noise-
inclusive variance) over this rank's 1 M test points.  With --gpus N (launched under torchrun,
one rank per GPU) the factor is computed on rank 0 and NCCL-broadcast, every rank owns its own
1 M-point shard (weak scaling) and there is no collective inside the timed region.

value     device-resident inputs / outputs (gpc_predict_dev), CUDA events on the core's stream
e2e       the reference-facing call (GPyMultiOutputWrapper.predict) with HOST buffers: pinned
          H2D of the test rows and D2H of mean and variance inside the timed region
roofline  the dominant kernel (k_vt: V = L^-1 K* as a DMMA contraction), algorithmic FLOPs =
          rows x N^2 per launch, against the cuBLAS DGEMM peak measured on this pool's B200
          (profiles/microbench/dgemm_peak_r01.json; MEASURED_PEAKS.json carries no FP64 figure)
cpu_baseline / --impl reference
          the NumPy/SciPy restatement of the reference's GPy/emukit arithmetic (oracle/, "port":
          GPy and emukit are not installable here) on the box's host cores, bounded sample
ig        secondary: RIG log-det information-gain evals/s (BASELINE configs[3], N = 4096, F = 3,
          65536 candidates x 32 points, G = 300 grid)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

entry.setup_path()

METRIC = "GP posterior mean+var test pts/sec"
MF2_PARAMS = np.array([4.0, 2.0, 3.0, 2.5, 1.0, 1.5, 2.0, 2.0, 0.8, 0.05, 0.02])   # var,l(3) x2, rho, noise x2
MF3_PARAMS = np.array([3.0, 2.5, 3.5, 3.0, 1.0, 1.5, 2.0, 2.0, 0.5, 1.0, 1.5, 1.5, 0.9, 1.1, 0.08, 0.04, 0.02])
DGEMM_PEAK_FILE = os.path.join(ROOT, "profiles", "microbench", "dgemm_peak_r01.json")
I8_PEAK_FILE = os.path.join(ROOT, "profiles", "microbench", "umma_i8_peak_r02.json")


# ------------------------------------------------------------------------------------------------
# synthetic workload (seeded)
# ------------------------------------------------------------------------------------------------
def wrbf_field(X):
    """Weighted-RBF scalar field of the reference's simulator (exploreSimSettings.py:74-101)."""
    WS = np.array([[0, 10.0], [0, 20.0]]); depth = 10.0
    p = np.array([[.7 * WS[0, 1], .7 * WS[1, 1], .5 * depth], [.3 * WS[0, 1], .2 * WS[1, 1], depth],
                  [.1 * WS[0, 1], .9 * WS[1, 1], depth], [.6 * WS[0, 1], .1 * WS[1, 1], .3 * depth],
                  [.1 * WS[0, 1], .1 * WS[1, 1], depth]])
    Lm, s, w = 10.0, 0.5, 0.5 * np.array([3.0, 2.0, 1.0])
    d2 = ((X[:, None, :] - p[None, :, :]) ** 2 * w[None, None, :]).sum(-1)
    return Lm * np.exp(-s * d2 / 10.0).sum(1)


def make_train(N, F, seed=0):
    rng = np.random.default_rng(seed)
    X = rng.uniform([0, 0, 0], [10, 20, 10], (N, 3))
    if F == 2:
        fid = (rng.uniform(size=N) < 0.2).astype(float)                 # 80 % low / 20 % high
        pert = np.where(fid[:, None] == 0, 0.5, 0.05)
    else:
        fid = rng.integers(0, F, N).astype(float)
        pert = np.choose(fid.astype(int), [0.5, 0.2, 0.05])[:, None]
    y = wrbf_field(X) + rng.normal(0, 0.125, N)
    Xobs = X + rng.normal(0, 1, (N, 3)) * pert                            # localisation error
    return np.hstack([Xobs, fid[:, None]]), y


def make_grid(n, fid, lo=0.0):
    ax = [np.linspace(lo, 10, n), np.linspace(lo, 20, n), np.linspace(lo, 10, n)]
    g = np.meshgrid(*ax)
    P = np.array([gi.ravel("F") for gi in g]).T
    # P is a transposed (Fortran-ordered) view and hstack keeps that order: force C rows (x, y, z, fid)
    return np.ascontiguousarray(np.hstack([P, np.full((P.shape[0], 1), float(fid))]))


def make_candidates(C, k, F, seed=1):
    rng = np.random.default_rng(seed)
    a = rng.uniform([0, 0, 0], [10, 20, 10], (C, 3))
    d = rng.normal(0, 1, (C, 3)); d *= (rng.uniform(0.5, 2.0, (C, 1)) / np.linalg.norm(d, axis=1, keepdims=True))
    t = np.linspace(0, 1, k)[None, :, None]
    pts = a[:, None, :] + t * d[:, None, :]
    var = np.linspace(0.05, 7.0, k)[None, :] * rng.uniform(0.5, 1.5, (C, 1))   # ramp crossing the thresholds
    fl = [0.25, 2.25, 6.25]
    fid = (var < fl[0]) * 2 + ((var > fl[0]) & (var < fl[1])) * 1
    rows = np.concatenate([pts, fid[:, :, None].astype(float)], axis=2).reshape(C * k, 4)
    return rows, np.arange(0, (C + 1) * k, k, dtype=np.int64)


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU arm (oracle port of the reference arithmetic)
# ------------------------------------------------------------------------------------------------
def cpu_posterior(N, M_sample, F, params, steps=1, warmup=0, tile=20000):
    """Times the NumPy/SciPy restatement on host cores: K* assembly, K*^T alpha, solve_triangular,
    diagonal variance -- the factorisation is done once outside (as on the GPU)."""
    from oracle import gp_oracle as go
    X4, y = make_train(N, F)
    t0 = time.perf_counter()
    gp = go.MFGP(X4, y, params, F=F)
    t_factor = time.perf_counter() - t0
    rng = np.random.default_rng(5)
    Xs4 = np.hstack([rng.uniform([0, 0, 0], [10, 20, 10], (M_sample, 3)), np.full((M_sample, 1), F - 1.0)])
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        for o in range(0, M_sample, tile):
            gp.predict(Xs4[o:o + tile])
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return M_sample * len(times) / sum(times), t_factor, sum(times) / len(times)


def blas_threads():
    """Threads the BLAS behind NumPy/SciPy actually uses (torchrun exports OMP_NUM_THREADS=1)."""
    try:
        from threadpoolctl import threadpool_info
        n = [p.get("num_threads", 1) for p in threadpool_info() if p.get("user_api") == "blas"]
        return max(n) if n else os.cpu_count()
    except Exception:
        return os.cpu_count()


def use_all_host_cores():
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=os.cpu_count())
    except Exception:
        pass


def cpu_nigp(X, y, noise_diag, M_sample, tile=10000):
    from oracle import gp_oracle as go
    ls, sf, sy = NIGP_HYP["ls"], NIGP_HYP["sigma_f"], NIGP_HYP["sigma_y"]
    t0 = time.perf_counter()
    K = go.SE_ARD_kernel(X, X, ls, sf)
    f = go.Factor(K + np.diag(sy ** 2 + noise_diag), y)
    t_factor = time.perf_counter() - t0
    rng = np.random.default_rng(5)
    Xs = rng.uniform([0, 0, 0], [10, 20, 10], (M_sample, 3))
    t0 = time.perf_counter()
    for o in range(0, M_sample, tile):
        Kxs = go.SE_ARD_kernel(Xs[o:o + tile], X, ls, sf)
        mean = Kxs @ f.alpha
        tmp = f.half_solve(Kxs.T)
        var = np.maximum(sf - np.sum(tmp * tmp, 0) + 1e-12, 1e-12)
    dt = time.perf_counter() - t0
    return M_sample / dt, t_factor


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    use_all_host_cores()
    N, F = args.n_train, 2
    M_sample = min(args.m_test, int(2e4 * 8192 / N) // 2)
    if args.workload == "nigp":
        X4, y = make_train(N, 3)
        nd = np.full(N, 0.01)
        ts = []
        for it in range(args.warmup + args.steps):
            p, t_factor = cpu_nigp(X4[:, :3], y, nd, M_sample)
            if it >= args.warmup:
                ts.append(M_sample / p)
        pts_s, t_step = M_sample * len(ts) / sum(ts), sum(ts) / len(ts)
    else:
        pts_s, t_factor, t_step = cpu_posterior(N, M_sample, F, MF2_PARAMS, steps=args.steps, warmup=args.warmup)
    cores = blas_threads()
    line = {"impl": "reference", "metric": METRIC, "value": pts_s, "unit": "pts/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args),
            "cpu_baseline": {"value": pts_s, "unit": "pts/s", "cores": cores, "kind": "port",
                             "sample": "%d of %d test points per step, N=%d, factor (%.2f s) outside the step; "
                                       "NumPy/SciPy restatement of the GPy/emukit arithmetic (neither is installable "
                                       "offline), BLAS threads = all host cores" % (M_sample, args.m_test, N, t_factor)},
            "e2e": {"value": pts_s, "unit": "pts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def oracle_parity(args, X4, y, noise_diag, Xs4_host, dmean, dvar, predict_api, nigp_mode, n=4096):
    """max_rel_* are NORMWISE (max|gpu - ref| / max(|ref|, prior variance), the 1e-9 criterion of the test-suite);
    max_elem_rel_var is the element-wise relative error of the variances (bounded below by the noise)."""
    from oracle import gp_oracle as go
    M = Xs4_host.shape[0]
    idx = np.unique(np.linspace(0, M - 1, n).astype(np.int64))
    Xq = np.ascontiguousarray(Xs4_host[idx])
    t0 = time.perf_counter()
    if nigp_mode:
        mu, var = go.nigp_predict(X4[:, :3], y, NIGP_HYP["ls"], NIGP_HYP["sigma_f"], NIGP_HYP["sigma_y"], noise_diag,
                                  Xq[:, :3], gram=False)
        scale, what = float(NIGP_HYP["sigma_f"]), "oracle.gp_oracle.nigp_predict (NIGP.py:269-333 restated, direct distances)"
        mu_h, var_h = predict_api(np.ascontiguousarray(Xq[:, :3]))
    else:
        ref = go.MFGP(X4, y, MF2_PARAMS, F=2, gram=False)
        mu, var = ref.predict(Xq)
        mu, var = mu[:, 0], var[:, 0]
        scale = float(MF2_PARAMS[0] * MF2_PARAMS[8] ** 2 + MF2_PARAMS[4])
        what = "oracle.gp_oracle.MFGP.predict (emukit AR1 arithmetic restated, direct distances)"
        mu_h, var_h = predict_api(Xq)
    t_cpu = time.perf_counter() - t0
    dm, dv = dmean.cpu().numpy()[idx], dvar.cpu().numpy()[idx]
    nrm = lambda a, b, s: float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), s))
    return {"n": int(idx.size), "against": what, "mode": args.mode,
            "max_rel_mean": max(nrm(dm, mu, 0.0), nrm(mu_h[:, 0], mu, 0.0)),
            "max_rel_var": max(nrm(dv, var, scale), nrm(var_h[:, 0], var, scale)),
            "max_elem_rel_var": float(max(np.max(np.abs(dv - var) / var), np.max(np.abs(var_h[:, 0] - var) / var))),
            "max_elem_rel_mean_floor_1e-3": float(np.max(np.abs(dm - mu) / np.maximum(np.abs(mu), 1e-3 * np.max(np.abs(mu))))),
            "min_var": float(var.min()), "definition": "max_rel_* normwise: max|gpu - ref| / max(max|ref|, prior variance); "
            "both the device-resident (value) and the host-buffer (e2e) results are checked", "oracle_s": t_cpu}


NIGP_HYP = dict(ls=np.array([2.0, 3.0, 2.5]), sigma_f=4.0, sigma_y=0.2, sigma_x=np.array([0.1, 0.1, 0.05]))


def apply_workload_defaults(args):
    """--workload mf2  = BASELINE configs[1] (the headline line);
       --workload nigp = BASELINE configs[2]: NIGP noisy-input GP, N = 8192, 2^21 test points per GPU
                         (16 M sharded over 8), per-point training noise from the posterior-mean gradients."""
    if args.workload == "nigp":
        args.n_train = args.n_train or 8192
        args.m_test = args.m_test or (1 << 21)
    else:
        args.n_train = args.n_train or 2048
        args.m_test = args.m_test or 1000000


def workload_config(args):
    cfg = _workload_config(args)
    if args.scaling == "strong":
        world = int(os.environ.get("WORLD_SIZE", "1"))
        cfg["workload"] += " -- STRONG scaling: --m-test = %d is the size of ONE test set split contiguously over %d GPU(s)" \
                           % (args.m_test, world)
        cfg["m_test_total"] = args.m_test
        cfg["m_test_per_gpu"] = args.m_test // world
    cfg["variance_path"] = ("int8: Ozaki splitting, 6 base-256 digits per operand, 21 exact digit GEMMs on the tcgen05 INT8 "
                            "tensor cores, FP64 results to ~1e-12 (parity tolerance 1e-9)") if args.mode == "int8" \
        else "fp64: DMMA (mma.sync.m8n8k4.f64)"
    return cfg


def _workload_config(args):
    if args.workload == "nigp":
        return {"workload": "configs[2]: NIGP noisy-input SE-ARD GP (sigma_x-derived per-point noise), N=%d train, "
                            "%d test points per GPU (a 16M-point set sharded over 8 GPUs), posterior mean + variance "
                            "(NIGP.predict semantics), fixed hypers" % (args.n_train, args.m_test),
                "n_train": args.n_train, "m_test_per_gpu": args.m_test, "fidelities": 1, "chunk_rows": args.chunk,
                "l2": "inputs larger than L2: each launch streams a %d MB K* chunk (> 126 MB L2)"
                      % (args.chunk * args.n_train * 8 // 2 ** 20)}
    return {"workload": "configs[1]: synthetic 2-fidelity AR1 (Kennedy-O'Hagan) GP, N=%d train, %d-point 3D test grid "
                        "at the top fidelity per GPU, posterior mean + noise-inclusive variance, fixed hypers"
                        % (args.n_train, args.m_test),
            "n_train": args.n_train, "m_test_per_gpu": args.m_test, "fidelities": 2, "rho": 0.8,
            "chunk_rows": args.chunk,
            "l2": "inputs larger than L2: each launch streams a %d MB K* chunk (> 126 MB L2); the 2 x %d MB factor "
                  "operands are meant to stay L2-resident" % (args.chunk * args.n_train * 8 // 2 ** 20,
                                                              args.n_train * args.n_train * 8 // 2 ** 20)}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import gpcore
    from gpcore import _lib as L
    from gpcore.emukit.multi_fidelity.kernels import LinearMultiFidelityKernel
    from gpcore.emukit.multi_fidelity.models import GPyLinearMultiFidelityModel
    from gpcore.emukit.model_wrappers.gpy_model_wrappers import GPyMultiOutputWrapper
    from gpcore.GPy.kern import RBF

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: gpcore has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ.pop("NCCL_DEBUG", None)     # NCCL prints its version banner to stdout at VERSION and above
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    gpcore.build()

    nigp_mode = args.workload == "nigp"
    N, M, F = args.n_train, args.m_test, (1 if nigp_mode else 2)
    X4, y = make_train(N, F if F > 1 else 3)
    if nigp_mode:
        X4[:, 3] = 0.0
    n = int(round(M ** (1.0 / 3)))
    strong = args.scaling == "strong"
    if strong:
        # ONE fixed test set of M_total points split contiguously over the ranks (configs[2]: 16 M points, N = 8192)
        from gpcore.sharding import shard_range
        M_total = M
        lo_s, hi_s = shard_range(M_total, rank, world)
        n = int(round(M_total ** (1.0 / 3)))
        if n ** 3 == M_total and M_total <= (1 << 24):
            Xs4_host = np.ascontiguousarray(make_grid(n, F - 1)[lo_s:hi_s])
        else:
            Xs4_host = np.hstack([np.random.default_rng(100).uniform([0, 0, 0], [10, 20, 10], (M_total, 3))[lo_s:hi_s],
                                  np.full((hi_s - lo_s, 1), F - 1.0)])
        M = hi_s - lo_s
    elif n ** 3 == M:
        # every rank owns a different 1 M-point grid (shifted by a sub-cell offset) -> weak scaling
        Xs4_host = make_grid(n, F - 1, lo=0.01 * rank)
    else:
        Xs4_host = np.hstack([np.random.default_rng(100 + rank).uniform([0, 0, 0], [10, 20, 10], (M, 3)),
                              np.full((M, 1), F - 1.0)])

    noise_diag = None
    if nigp_mode:
        # per-point input-noise variance v_i = sum_d (d mean / d x_d)^2 sigma_x_d^2 (NIGP.py:222,251-252)
        from gpcore import nigp as gnigp
        _, g = gnigp.compute_post_mean_and_gradients(X4[:, :3], y, NIGP_HYP["ls"], NIGP_HYP["sigma_f"],
                                                     NIGP_HYP["sigma_y"], device=local)
        noise_diag = np.sum(g ** 2 * NIGP_HYP["sigma_x"][None, :] ** 2, axis=1)
        core = gpcore.GPCore(L.KIND_NIGP, 1, local)
        core.set_chunk(args.chunk)
        core.set_hypers(np.concatenate([NIGP_HYP["ls"], [NIGP_HYP["sigma_f"], NIGP_HYP["sigma_y"]]]), 0.0)
        core.set_data(X4, y, noise_diag)
    else:
        core = gpcore.GPCore(L.KIND_MF_AR1_RBF, F, local)
        core.set_chunk(args.chunk)
        core.set_hypers(MF2_PARAMS, 1e-8)
        core.set_data(X4, y)
    mode = L.MODE_INT8 if args.mode == "int8" else L.MODE_FP64
    core.set_mode(mode)
    from gpcore.sharding import replicate_factor, warm_up
    if world > 1:
        warm_up()            # communicator set-up (~1 s at 8 ranks) is not part of any factor broadcast
    core.factor()            # first call: buffer allocation + first-touch; the timed replication below runs warm

    stream = torch.cuda.ExternalStream(core.stream())
    assert Xs4_host.flags["C_CONTIGUOUS"]
    dXs = torch.from_numpy(Xs4_host).cuda().contiguous()
    dmean = torch.empty(M, dtype=torch.float64, device="cuda")
    dvar = torch.empty(M, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    flags = L.NIGP_FLOOR if nigp_mode else (L.INCLUDE_NOISE | L.CLIP_DIAG)

    def step_dev():
        core.predict_dev(dXs.data_ptr(), M, dmean.data_ptr(), dvar.data_ptr(), flags)

    # ---- the serial part: (re)factor -- redundantly on every rank up to N = 4096, else on rank 0 with the lower
    # block-triangles of L and L^-1 NCCL-broadcast -- and the time to the FIRST result (factor + replication + one
    # launch batch of this rank's points), max over ranks.  Warm buffers, outside the throughput region. ----------
    m_first = min(M, args.chunk)
    core.predict_dev(dXs.data_ptr(), m_first, dmean.data_ptr(), dvar.data_ptr(), flags)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    repl = replicate_factor(core, 0)
    core.predict_dev(dXs.data_ptr(), m_first, dmean.data_ptr(), dvar.data_ptr(), flags)
    torch.cuda.synchronize()
    ttfr_ms = 1e3 * (time.perf_counter() - t0)
    if repl["mode"].startswith("broadcast"):
        # the same replication once more (staging buffers of the bands now cached by the allocator): steady state
        dist.barrier()
        again = replicate_factor(core, 0)
        repl["steady_ms"], repl["steady_gbps"], repl["steady_factor_ms"] = again["ms"], again["gbps"], again["factor_ms"]
    if dist is not None:
        t = torch.tensor([ttfr_ms, repl["ms"], repl["factor_ms"]], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ttfr_ms, repl["ms"], repl["factor_ms"] = float(t[0]), float(t[1]), float(t[2])
        if repl["ms"] > 0:
            repl["gbps"] = repl["bytes"] / (repl["ms"] * 1e-3) / 1e9
    t_factor, t_bcast = 1e-3 * repl["factor_ms"], 1e-3 * repl["ms"]

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t[0])
        return ms

    # ---- value: device-resident -----------------------------------------------------------------
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    core.enable_hot_timing(True)
    for _ in range(args.warmup):
        step_dev()
    barrier()
    core.hot_kernel_time(reset=True)
    l0 = core.launch_count()
    ms = timed(step_dev, args.steps, 0)
    launches = core.launch_count() - l0
    hot_ms, hot_n, hot_flops = core.hot_kernel_time(reset=True)
    core.enable_hot_timing(False)
    clk = clocks.stop() if rank == 0 else None
    pts_all_ranks = (M_total if strong else world * M)
    value = pts_all_ranks * args.steps / (ms * 1e-3)

    # ---- e2e: reference-facing API, ORDINARY (pageable) NumPy arrays in and out, as a reference script passes them;
    # the library stages them through its own page-locked ring.  Copies inside the timed region.  The same call with
    # an input array the caller pinned is timed next to it (e2e.pinned_input_value) -------------------------------
    out = {}
    if nigp_mode:
        from gpcore.nigp import NIGP
        wrap = NIGP(verbose=False, device=local)
        wrap.lengthscales_, wrap.sigma_f_, wrap.sigma_y_ = NIGP_HYP["ls"], NIGP_HYP["sigma_f"], NIGP_HYP["sigma_y"]
        wrap.sigma_x_, wrap.X_train_, wrap.y_train_, wrap.noise_diag_train_ = NIGP_HYP["sigma_x"], X4[:, :3], y, noise_diag
        wrap._factor().set_chunk(args.chunk)
        wrap._factor().set_mode(mode)
        Xs_plain = np.ascontiguousarray(Xs4_host[:, :3])
        pinned = torch.from_numpy(Xs_plain).pin_memory()
        Xs_pinned = pinned.numpy()
        api = "gpcore.nigp.NIGP.predict(Xs) -> gpc_predict (pageable NumPy arrays, staged through the library's pinned ring)"
        h2d = M * 32

        def predict_api(Xq):
            mu, var = wrap.predict(Xq)
            return mu[:, None], var[:, None]
    else:
        kern = LinearMultiFidelityKernel([RBF(3, ARD=True), RBF(3, ARD=True)])
        model = GPyLinearMultiFidelityModel(X4, y[:, None], kern, n_fidelities=F, device=local)
        model.param_array[:] = MF2_PARAMS
        wrap = GPyMultiOutputWrapper(model, F, n_optimization_restarts=1)
        model._ensure_factor().set_chunk(args.chunk)
        model._ensure_factor().set_mode(mode)
        Xs_plain = Xs4_host
        pinned = torch.from_numpy(Xs4_host).pin_memory()
        Xs_pinned = pinned.numpy()
        api = ("gpcore.emukit GPyMultiOutputWrapper.predict(X4) -> gpc_predict (pageable NumPy arrays, staged through "
               "the library's pinned ring)")
        h2d = M * 32
        predict_api = wrap.predict

    def time_e2e(Xq):
        def step_e2e():
            mu, var = predict_api(Xq)
            out["chk"] = float(mu[0, 0]) + float(var[-1, 0])

        for _ in range(args.warmup):
            step_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_e2e()
        barrier()
        dt = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t[0])
        return pts_all_ranks * args.steps / dt

    e2e = time_e2e(Xs_plain)
    e2e_pinned = time_e2e(Xs_pinned)
    # parity spot check of the e2e result against the device-resident result
    mu, var = predict_api(Xs_pinned[:4096])
    dm, dv = dmean[:4096].cpu().numpy(), dvar[:4096].cpu().numpy()
    err_m, err_v = float(np.max(np.abs(mu[:, 0] - dm))), float(np.max(np.abs(var[:, 0] - dv)))
    if not (err_m <= 1e-9 * max(1.0, float(np.max(np.abs(dm)))) and err_v <= 1e-9 * MF2_PARAMS[0]):
        raise SystemExit("e2e and device-resident results disagree: mean %.3e var %.3e" % (err_m, err_v))

    # ---- parity of what was just timed, against the CPU oracle (the checker, never the thing measured): a strided
    # 4096-point sample of this rank's test set, device-resident (value) and host-buffer (e2e) results ------------
    parity = None
    if rank == 0 and args.parity:
        parity = oracle_parity(args, X4, y, noise_diag, Xs4_host, dmean, dvar, predict_api, nigp_mode)
        if not (parity["max_rel_mean"] <= 1e-9 and parity["max_rel_var"] <= 1e-9):
            raise SystemExit("parity against the oracle failed: %s" % json.dumps(parity))

    # configs[2] also asks for the Xs_input_noise path (NIGP.py:304-324: + sum_d (d mean / d x_d)^2 sigma_x_d^2,
    # needs the mean gradients): one untimed and one timed call through the reference-facing API
    noisy = None
    if nigp_mode and dist is None:
        wrap.predict(Xs_pinned, Xs_input_noise=NIGP_HYP["sigma_x"])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        mu_n, var_n = wrap.predict(Xs_pinned, Xs_input_noise=NIGP_HYP["sigma_x"])
        dtn = time.perf_counter() - t0
        noisy = {"value": M / dtn, "unit": "pts/s", "api": "gpcore.nigp.NIGP.predict(Xs, Xs_input_noise=sigma_x) (host buffers)",
                 "finite": bool(np.all(np.isfinite(var_n)) and np.all(np.isfinite(mu_n)))}

    # ---- RIG information gain (the metric's second half): candidates sharded over ALL ranks, best node by NCCL -----
    ig_line = None
    if args.ig and not nigp_mode:
        ig_line = bench_ig(args, gpcore, L, torch, dist, rank, world, local)

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel -----------------------------------------------------------
    dgemm_peak, dgemm_src = 35.46, "fallback: cuBLAS DGEMM 8192^3 measured on this pool's B200 in round 1"
    try:
        pk = json.load(open(DGEMM_PEAK_FILE))
        dgemm_peak = float(pk["dgemm_8192_tflops"])
        dgemm_src = "measured: cuBLAS DGEMM 8192^3 on this pool's B200 (profiles/microbench/dgemm_peak_r01.json); " \
                    "MEASURED_PEAKS.json has no FP64 entry"
    except Exception:
        pass
    traffic = None
    try:   # DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture (same config only)
        tr = json.load(open(os.path.join(ROOT, "profiles", "r02" if args.mode == "int8" else "r01",
                                         "k_vt_i8_traffic.json" if args.mode == "int8" else "k_vt_traffic.json")))
        if tr["n_train"] == N and tr["chunk_rows"] == args.chunk and not nigp_mode:
            traffic = tr["dram_bytes_per_launch"]
    except Exception:
        pass
    rows_per_launch = M * args.steps / max(hot_n, 1)
    alg_flops_per_launch = rows_per_launch * float(N) * float(N)     # FP64 algorithmic work: N^2 per point
    avg_ms = hot_ms / max(hot_n, 1)
    fp64_equiv = alg_flops_per_launch / (avg_ms * 1e-3) / 1e12
    if args.mode == "int8":
        # 21 exact int8 digit GEMMs stand in for the FP64 contraction: report the kernel against the INT8 tensor peak
        # the kernel is timed inside a long step that runs at the 1000 W power cap (profiles/r02/phase_power_r02.json):
        # the roofline denominator is the SUSTAINED tcgen05 INT8 rate measured under the same cap; the burst rate
        # (one isolated 2 ms launch) is listed next to it
        i8_burst, i8_peak, i8_src = int8_peaks()
        achieved = 21.0 * alg_flops_per_launch / (avg_ms * 1e-3) / 1e12
        roofline = {"bound": "tensor", "kernel": "k_vt_i8<OUT_SUMSQ> (V = L^-1 K* as 21 exact INT8 digit GEMMs, tcgen05 + TMEM, "
                                                 "fused recombination and sum of squares)",
                    "achieved": achieved, "peak": i8_peak, "unit": "TOP/s (int8)", "frac": achieved / i8_peak, "traffic": traffic,
                    "peak_burst": i8_burst, "frac_of_burst": achieved / i8_burst,
                    "peak_source": i8_src + " (tcgen05.mma.kind::i8 M=128 N=256, random digit bytes; sustained = back-to-back "
                                   "launches at the board's 1000 W power cap, which is where this step runs -- clocks.reasons "
                                   "shows sw_power_cap; MEASURED_PEAKS.json has no INT8 entry)",
                    "launches": hot_n, "avg_launch_ms": avg_ms,
                    "algorithmic_ops_per_launch": 21.0 * alg_flops_per_launch,
                    "executed_ops_per_launch": 21.0 * hot_flops / max(hot_n, 1),
                    "fp64_equivalent_tflops": fp64_equiv, "fp64_dgemm_peak_tflops": dgemm_peak,
                    "fp64_equivalent_over_dgemm_peak": fp64_equiv / dgemm_peak,
                    "share_of_step": hot_ms / ms}
    else:
        roofline = {"bound": "tensor", "kernel": "k_vt<false,true> (V = L^-1 K*, FP64 DMMA, fused sum of squares)",
                    "achieved": fp64_equiv, "peak": dgemm_peak, "unit": "TFLOP/s", "frac": fp64_equiv / dgemm_peak,
                    "traffic": traffic, "peak_source": dgemm_src, "launches": hot_n, "avg_launch_ms": avg_ms,
                    "algorithmic_flops_per_launch": alg_flops_per_launch,
                    "executed_flops_per_launch": hot_flops / max(hot_n, 1),
                    "share_of_step": hot_ms / ms}
    # the same workload on the other contraction path, a short run for context (single-GPU runs only:
    # this block runs on rank 0 alone and must not touch a collective)
    other = {}
    if dist is None:
        try:
            omode = L.MODE_FP64 if args.mode == "int8" else L.MODE_INT8
            core.set_mode(omode)
            for _ in range(2):
                step_dev()
            torch.cuda.synchronize()
            oms = timed(step_dev, 2, 0)
            other = {"mode": "fp64" if args.mode == "int8" else "int8", "value": M * 2 / (oms * 1e-3), "ms_per_step": oms / 2}
            # the optional FP32-tolerance mode of the north star (4 of the 6 digits: 10 digit GEMMs), context only
            core.set_mode(L.MODE_INT8_F32)
            for _ in range(2):
                step_dev()
            torch.cuda.synchronize()
            fms = timed(step_dev, 3, 0)
            other["int8_f32_tolerance_mode"] = {"value": M * 3 / (fms * 1e-3), "ms_per_step": fms / 3,
                                                "note": "GPC_MODE_INT8_F32: variances to ~1e-7 relative (north_star allows 1e-4 "
                                                        "for an optional FP32 mode); NOT the headline metric"}
            # five of the six digit levels (15 digit GEMMs): the error-budgeted variant VERDICT r1 asked to evaluate.
            # Timed AND checked against the oracle here; not the default because its margin to 1e-9 is a factor of two
            core.set_mode(L.MODE_INT8_L5)
            for _ in range(2):
                step_dev()
            torch.cuda.synchronize()
            lms = timed(step_dev, 3, 0)
            l5 = {"value": M * 3 / (lms * 1e-3), "ms_per_step": lms / 3}
            if args.parity and not nigp_mode:
                from oracle import gp_oracle as go
                idx = np.unique(np.linspace(0, M - 1, 4096).astype(np.int64))
                refp = go.MFGP(X4, y, MF2_PARAMS, F=2, gram=False)
                _, vr = refp.predict(np.ascontiguousarray(Xs4_host[idx]))
                dv5 = dvar.cpu().numpy()[idx]
                sc5 = float(MF2_PARAMS[0] * MF2_PARAMS[8] ** 2 + MF2_PARAMS[4])
                l5["max_rel_var"] = float(np.max(np.abs(dv5 - vr[:, 0])) / max(float(np.max(vr)), sc5))
                l5["max_elem_rel_var"] = float(np.max(np.abs(dv5 - vr[:, 0]) / vr[:, 0]))
            l5["note"] = "GPC_MODE_INT8_L5: opt-in; normwise margin to 1e-9 is ~2x, information gain is outside 1e-9 in this mode"
            other["int8_five_levels"] = l5
            core.set_mode(mode)
            step_dev()               # leave the default mode's results in the output buffers
            torch.cuda.synchronize()
        except Exception as exc:   # context only
            other = {"error": str(exc)}

    # ---- posterior MEAN only (K* alpha fused into the covariance assembly, no contraction): context for the
    # north_star's 1e9 points/s figure, which is a mean-only / FP64-ALU-bound rate -----------------
    mean_only = None
    if dist is None:
        def step_mean():
            core.predict_dev(dXs.data_ptr(), M, dmean.data_ptr(), 0, flags | L.MEAN_ONLY)
        for _ in range(2):
            step_mean()
        torch.cuda.synchronize()
        mms = timed(step_mean, 3, 0)
        terms = 1.0 if nigp_mode else 1.2      # AR1 terms per (test, train) pair of this workload
        terms_grid = 1.0 if nigp_mode else 2.0  # the grid path contracts every term over all training points
        # the same mean on the benchmark's test set AS A TENSOR GRID (it is one): gpc_predict_grid_mean, device output
        side = int(round(M ** (1.0 / 3)))
        grid_line = None
        if side ** 3 == M:
            axs = [np.linspace(0.0, 10, side), np.linspace(0.0, 20, side), np.linspace(0.0, 10, side)]
            gm = torch.empty(M, dtype=torch.float64, device="cuda")
            core.predict_grid_mean_dev(axs[0], axs[1], axs[2], F - 1, gm.data_ptr())
            torch.cuda.synchronize()
            gms = timed(lambda: core.predict_grid_mean_dev(axs[0], axs[1], axs[2], F - 1, gm.data_ptr()), 3, 0)
            grid_line = {"value": M * 3 / (gms * 1e-3), "unit": "pts/s", "ms_per_step": gms / 3,
                         "dgemm_tflops": 2.0 * M * float(N) * terms_grid * 3 / (gms * 1e-3) / 1e12,
                         "api": "gpc_predict_grid_mean_dev (separable cross-covariance, contraction over the training "
                                "index as FP64 DMMA GEMMs)"}
        mean_only = {"value": M * 3 / (mms * 1e-3), "unit": "pts/s", "ms_per_step": mms / 3, "tensor_grid": grid_line,
                     "kernel_evals_per_s": M * 3 * float(N) * terms / (mms * 1e-3),
                     "exp_peak_per_s": 784e9, "note": "N kernel evaluations + 2N flop per point; library exp(double) "
                     "peaks at 784 G/s on this B200 (profiles/microbench/fp64_peaks_r01.txt)"}

    # ---- the factorisation (once per hyper-parameter / data change): assembly + Cholesky + explicit L^-1 +
    # alpha + log-det, warm buffers, best of 3 -------------------------------------------------------------
    factor = None
    if dist is None:
        tf = []
        for _ in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            core.factor()
            tf.append(time.perf_counter() - t0)
        tfl = 2.0 * float(N) ** 3 / 3.0 / min(tf) / 1e12
        # the same call with the triangular inverse switched off (profiling switch): what a potrf-only library call covers
        os.environ["GPC_FACTOR_PHASE"] = "chol"
        tc = []
        try:
            for _ in range(4):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                try:
                    core.factor()
                except Exception:
                    pass
                tc.append(time.perf_counter() - t0)
        finally:
            os.environ.pop("GPC_FACTOR_PHASE", None)
        core.factor()                 # restore a valid factor
        factor = {"ms": 1e3 * min(tf), "tflops": tfl, "frac_of_dgemm_peak": tfl / dgemm_peak,
                  "ms_chol_only": 1e3 * min(tc[1:]),
                  "flops": "N^3/3 Cholesky + N^3/3 triangular inverse (assembly, alpha, log-det inside the time); ms_chol_only: "
                           "the same call without the inverse (cuSOLVER potrf at N = 2048 on this pool: 0.97 ms, Cholesky only)"}

    # ---- the factorisation at N = 8192 (configs[2]'s size), both ways of counting: Cholesky + explicit L^-1 as the
    # product computes it (2 N^3 / 3 flop) and the Cholesky alone (GPC_FACTOR_PHASE=chol, a profiling switch: N^3 / 3)
    factor_8192 = None
    if dist is None and args.factor8192:
        factor_8192 = bench_factor(gpcore, L, torch, local, 8192, dgemm_peak)

    # ---- CPU baseline (bounded sample) -------------------------------------------------------------
    cpu = None
    if world == 1:
        M_cpu = min(M, int(2e4 * 8192 / N) // 2)
        if nigp_mode:
            cpu_pts, cpu_factor = cpu_nigp(X4[:, :3], y, noise_diag, M_cpu)
            what = "the oracle restatement of NIGP.predict (diagonal only; the reference builds the M x M matrix)"
        else:
            cpu_pts, cpu_factor, _ = cpu_posterior(N, M_cpu, F, MF2_PARAMS, steps=1, warmup=0)
            what = "NumPy/SciPy restatement of the GPy/emukit arithmetic"
        cpu = {"value": cpu_pts, "unit": "pts/s", "cores": blas_threads(), "kind": "port",
               "sample": "%d test points (tile of the %d-point workload), N=%d; factor %.2f s outside the sample; %s"
                         % (M_cpu, M, N, cpu_factor, what)}

    line = {"metric": METRIC, "value": value, "unit": "pts/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if strong else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args),
            "e2e": {"value": e2e, "unit": "pts/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(M * 16),
                    "api": api, "host_buffers": "pageable (plain numpy.ndarray) in and out", "pinned_input_value": e2e_pinned},
            "gpu_launches": int(launches), "clocks": clk, "roofline": roofline, "cpu_baseline": cpu,
            "mode": args.mode, "other_mode": other,
            "factor_ms": 1e3 * t_factor, "factor_broadcast_ms": 1e3 * t_bcast, "factor_replication": repl,
            "time_to_first_result_ms": ttfr_ms, "factor": factor, "factor_8192": factor_8192,
            "mean_only": mean_only, "noisy_input": noisy, "parity": parity}

    line["ig"] = ig_line
    emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def bench_factor(gpcore, L, torch, local, N, dgemm_peak):
    X4, y = make_train(N, 3, seed=8)
    X4[:, 3] = 0.0
    core = gpcore.GPCore(L.KIND_SF_RBF, 1, local)
    core.set_hypers(np.array([4.0, 2.0, 3.0, 2.5, 0.05]), 1e-8)
    core.set_data(X4, y)
    out = {}
    for phase in ("full", "chol"):
        if phase == "chol":
            os.environ["GPC_FACTOR_PHASE"] = "chol"
        try:
            core.factor()
        except Exception:
            pass                      # chol-only leaves no valid inverse: alpha / log-det of that pass are garbage
        ts = []
        for _ in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            try:
                core.factor()
            except Exception:
                pass
            ts.append(time.perf_counter() - t0)
        out[phase] = min(ts)
        os.environ.pop("GPC_FACTOR_PHASE", None)
    core.close()
    n3 = float(N) ** 3 / 3.0
    return {"n_train": N, "ms": 1e3 * out["full"], "ms_chol_only": 1e3 * out["chol"],
            "tflops_with_inverse": 2.0 * n3 / out["full"] / 1e12, "frac_with_inverse": 2.0 * n3 / out["full"] / 1e12 / dgemm_peak,
            "tflops_chol_flops_over_full_time": n3 / out["full"] / 1e12,
            "tflops_chol_only": n3 / out["chol"] / 1e12, "frac_chol_only": n3 / out["chol"] / 1e12 / dgemm_peak,
            "note": "assembly, alpha (two triangular solves with one refinement step each) and the log-det are inside both "
                    "times; ms_chol_only = the same call with the triangular inverse switched off (GPC_FACTOR_PHASE=chol)"}


def int8_peaks():
    """tcgen05 INT8 peaks measured on this pool's B200 by profiles/microbench/umma_i8_sustained.cu (random digit bytes):
    burst (one 2 ms launch) and sustained (back-to-back launches, the 1000 W power cap sets the clock)."""
    try:
        pk = json.load(open(I8_PEAK_FILE))
        return float(pk["burst_tops"]), float(pk["sustained_tops"]), "measured: profiles/microbench/umma_i8_peak_r02.json"
    except Exception:
        return 4341.7, 3774.4, "fallback: the figures profiles/microbench/umma_i8_sustained.cu measured in round 2"


def cpu_ig_baseline(N, F, k, grid4, X4, y, rows, offs, n_literal=3, n_schur=16):
    """The reference's way (calculatePathInfoEmuBatch, PhysicalExperimentCode/GraceRIGV3.py:599-618: one full refit of
    the (N + k)-point model and two G x G determinants per candidate) on a bounded sample, and the Schur-complement
    port of the same quantity, both on the host cores."""
    from oracle import gp_oracle as go
    t0 = time.perf_counter()
    ref = go.MFGP(X4, y, MF3_PARAMS, F=F)
    t_fit = time.perf_counter() - t0
    t0 = time.perf_counter()
    lit = [go.ig_logdet_refit(ref, grid4, rows[offs[c]:offs[c + 1]], clip_cov=1e-10) for c in range(n_literal)]
    t_lit = time.perf_counter() - t0
    _, _, _, noise = go.split_mf_params(MF3_PARAMS, F)
    t0 = time.perf_counter()
    sch = [go.ig_logdet_schur(ref, grid4, rows[offs[c]:offs[c + 1]], noise[2], noise[rows[offs[c]:offs[c + 1], 3].astype(int)])
           for c in range(n_schur)]
    t_sch = time.perf_counter() - t0
    return {"value": n_literal / t_lit, "unit": "evals/s", "cores": blas_threads(), "kind": "port",
            "sample": "%d of the candidates by the literal refit loop (one O((N + k)^3) refit + two 300 x 300 determinants "
                      "per candidate, as the reference does; model fit %.1f s outside the sample), N=%d, k=%d"
                      % (n_literal, t_fit, N, k),
            "schur_port_value": n_schur / t_sch, "schur_port_sample": "%d candidates by the k x k determinant-lemma form" % n_schur,
            "finite": bool(np.all(np.isfinite(lit)) and np.all(np.isfinite(sch)))}


def bench_ig(args, gpcore, L, torch, dist, rank, world, local):
    """RIG info-gain evaluations per second (BASELINE metric, second half; configs[3]): N = 4096 three-fidelity MF-GP,
    65536 candidate nodes x 32 points, G = 300 grid.  The factor is replicated (every rank factors: N <= 4096), the
    candidates are sharded contiguously over the ranks, every rank scores its share, and one 16-byte all-gather picks the
    best node (gpcore.sharding.reduce_best) -- inside the timed region, because it is part of the path it replaces
    (the serial scan GraceRIGV3.py:1072-1189).  Host buffers in, scores + argmax out (the C ABI call a planner makes)."""
    from gpcore.sharding import reduce_best, shard_candidates
    N, F, C, k, G = 4096, 3, args.ig_candidates, 32, 300
    X4, y = make_train(N, F, seed=3)
    core = gpcore.GPCore(L.KIND_MF_AR1_RBF, F, local)
    core.set_hypers(MF3_PARAMS, 1e-8)
    core.set_data(X4, y)
    core.factor()
    g = np.meshgrid(np.linspace(0, 10, 10), np.linspace(0, 20, 6), np.linspace(0, 10, 5))
    grid4 = np.ascontiguousarray(np.hstack([np.array([gi.ravel("F") for gi in g]).T, 2 * np.ones((G, 1))]))
    rows, offs = make_candidates(C, k, F)                 # the same seeded set on every rank
    c_lo, c_hi, loffs, r_lo, r_hi = shard_candidates(offs, rank, world)
    lrows = np.ascontiguousarray(rows[r_lo:r_hi])
    sig_n = float(MF3_PARAMS[-1])
    ops = {"logdet": lambda: core.ig_logdet(grid4, lrows, loffs)[::2],
           "seq": lambda: core.ig_seq(lrows, loffs, sig_n, pred_fid=0),
           "logdet_clip": lambda: core.ig_logdet(grid4, lrows, loffs, clip=True)[::2]}

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    res, scores = {}, {}
    core.enable_hot_timing(True)
    for name, fn in ops.items():
        fn()                                              # warm-up at full size (buffers, first touch)
        barrier()
        core.hot_kernel_time(reset=True)
        t0 = time.perf_counter()
        I, lbest = fn()
        bv, bi = (float(I[lbest]), int(lbest) + c_lo) if dist is None else \
            reduce_best(I[lbest] if lbest >= 0 else 0.0, lbest, c_lo)
        dt = time.perf_counter() - t0
        dev_ms = core.last_call_device_ms()
        hot_ms, hot_n, hot_fl = core.hot_kernel_time(reset=True)
        if dist is not None:
            t = torch.tensor([dt, dev_ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt, dev_ms = float(t[0]), float(t[1])
        res[name] = dict(dt=dt, dev_ms=dev_ms, hot_ms=hot_ms, hot_n=hot_n, hot_fl=hot_fl, best=bi, best_value=bv,
                         finite=bool(np.all(np.isfinite(I))))
        scores[name] = I
    core.enable_hot_timing(False)
    # the sharded winner must be the single-rank argmax over ALL candidates (rank 0 scores the full set, untimed)
    verified = None
    if rank == 0:
        if world > 1:
            If, _, bf = core.ig_logdet(grid4, rows, offs)
            Isf, bsf = core.ig_seq(rows, offs, sig_n, pred_fid=0)
            verified = bool(res["logdet"]["best"] == int(bf) and res["logdet"]["best_value"] == float(If[bf])
                            and res["seq"]["best"] == int(bsf) and res["seq"]["best_value"] == float(Isf[bsf]))
        else:
            verified = bool(res["logdet"]["best"] == int(np.argmax(scores["logdet"])))
    core.close()
    if rank != 0:
        return None
    burst, sustained, src = int8_peaks()
    ld = res["logdet"]
    # SURVEY 8d, per candidate: k N^2 (V = L^-1 K*) + 2 k G N (cross term with the grid's V) + k^2 N (Gram); each FP64
    # multiply-add pair is 21 exact INT8 digit products on the tensor cores
    alg = float(c_hi - c_lo) * (k * float(N) * N + 2.0 * k * G * N + float(k) * k * N)
    ach = 21.0 * alg / (ld["hot_ms"] * 1e-3) / 1e12
    line = {"metric": "RIG info-gain evals/sec", "unit": "evals/s", "n_gpus": world, "scaling": "strong",
            "value": C / (ld["dev_ms"] * 1e-3), "operator": "log-det IG on the fixed grid (gpc_ig_logdet: calcPathInfoSFBatch / "
            "calculatePathInfoEmuBatch without emukit's clip), device time by CUDA events (staged uploads included), max over ranks",
            "e2e": {"value": C / ld["dt"], "unit": "evals/s", "h2d_bytes_per_step": int(lrows.nbytes + loffs.nbytes + grid4.nbytes),
                    "d2h_bytes_per_step": int(8 * (c_hi - c_lo) + 8),
                    "api": "GPCore.ig_logdet -> gpc_ig_logdet_ex (pageable host arrays in, scores + argmax out) + reduce_best"},
            "config": {"workload": "configs[3]: RIG info-gain scoring of %d candidate tree nodes x %d points against an N=%d "
                                   "three-fidelity MF-GP, %d-point grid, candidates sharded over %d GPU(s)" % (C, k, N, G, world),
                       "n_train": N, "fidelities": F, "candidates": C, "points_per_candidate": k, "grid": G},
            "operators": {nm: {"value": C / (r["dev_ms"] * 1e-3), "e2e": C / r["dt"], "best": r["best"], "finite": r["finite"]}
                          for nm, r in res.items()},
            "sharded_best_equals_single_rank_argmax": verified,
            "roofline": {"bound": "tensor", "kernel": "k_vt_i8<OUT_DIGITS> + k_vt_i8<OUT_F64, FULLK> (V = L^-1 K* emitted as a digit "
                         "image, diagonal-tile Gram blocks and cross products with the grid's V from it), rank 0's launches",
                         "achieved": ach, "peak": sustained, "unit": "TOP/s (int8)", "frac": ach / sustained,
                         "peak_burst": burst, "frac_of_burst": ach / burst, "peak_source": src + " (sustained: the kernels run "
                         "inside a long power-capped step)", "launches": ld["hot_n"], "avg_launch_ms": ld["hot_ms"] / max(ld["hot_n"], 1),
                         "algorithmic_ops": 21.0 * alg, "executed_ops": 21.0 * ld["hot_fl"], "share_of_call": ld["hot_ms"] / ld["dev_ms"],
                         "traffic": None},
            "cpu_baseline": cpu_ig_baseline(N, F, k, grid4, X4, y, rows, offs) if world == 1 else None}
    return line


_STDOUT_FD = None


def quiet_stdout():
    """Everything that native libraries write to fd 1 during the run (NCCL's version banner ...) goes to stderr:
    stdout carries exactly the one JSON line."""
    global _STDOUT_FD
    sys.stdout.flush()
    _STDOUT_FD = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    if _STDOUT_FD is not None:
        os.dup2(_STDOUT_FD, 1)
    print(json.dumps(line), flush=True)
    if _STDOUT_FD is not None:
        os.dup2(2, 1)


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="mf2", choices=["mf2", "nigp"])
    ap.add_argument("--mode", default="int8", choices=["int8", "fp64"],
                    help="variance contraction: int8 = Ozaki digits on the tcgen05 INT8 tensor cores (default), "
                         "fp64 = DMMA")
    ap.add_argument("--n-train", type=int, default=0)
    ap.add_argument("--m-test", type=int, default=0)
    ap.add_argument("--chunk", type=int, default=65536)
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --m-test points PER GPU (default); strong: ONE set of --m-test points split over the GPUs")
    ap.add_argument("--parity", type=int, default=1, help="check the timed results against the CPU oracle (4096 points)")
    ap.add_argument("--factor8192", type=int, default=1, help="also time the factorisation at N = 8192")
    ap.add_argument("--ig", type=int, default=1)
    ap.add_argument("--ig-candidates", type=int, default=65536)
    args = ap.parse_args()
    apply_workload_defaults(args)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
