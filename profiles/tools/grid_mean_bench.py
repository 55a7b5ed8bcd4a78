"""Posterior mean on a tensor grid: gpc_predict_grid_mean (GEMMs on the FP64 tensor cores) against the general
mean-only predict (one kernel evaluation per (test, train) pair), host buffers in / out."""
import sys, os, time, numpy as np
sys.path.insert(0, os.getcwd())
import __graft_entry__ as e; e.setup_path()
import gpcore, bench, torch
from gpcore import _lib as L
for N, F, side, params in ((2048, 2, 100, bench.MF2_PARAMS), (8192, 1, 128, np.array([4.0, 2.0, 3.0, 2.5, 0.04])), (8192, 1, 256, np.array([4.0, 2.0, 3.0, 2.5, 0.04]))):
    X4, y = bench.make_train(N, F if F > 1 else 3)
    if F == 1:
        X4[:, 3] = 0.0
    core = gpcore.GPCore(L.KIND_SF_RBF if F == 1 else L.KIND_MF_AR1_RBF, F, 0)
    core.set_hypers(params, 1e-8); core.set_data(X4, y); core.factor()
    ax, ay, az = np.linspace(0, 10, side), np.linspace(0, 20, side), np.linspace(0, 10, side)
    core.predict_grid_mean(ax, ay, az, fid=F - 1)
    torch.cuda.synchronize()
    t0 = time.perf_counter(); m = core.predict_grid_mean(ax, ay, az, fid=F - 1); dt = time.perf_counter() - t0
    M = side ** 3
    dm = torch.empty(M, dtype=torch.float64, device="cuda")
    stream = torch.cuda.ExternalStream(core.stream())
    core.predict_grid_mean_dev(ax, ay, az, F - 1, dm.data_ptr())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream); core.predict_grid_mean_dev(ax, ay, az, F - 1, dm.data_ptr()); e1.record(stream)
    torch.cuda.synchronize()
    dev_ms = e0.elapsed_time(e1)
    assert np.array_equal(dm.cpu().numpy(), m.ravel())
    line = "N=%d F=%d grid %d^3: grid mean host in/out %.3g pts/s (%.1f ms); device-resident output %.3g pts/s (%.2f ms, %.1f TFLOP/s of 2MN per term)" % (
        N, F, side, M / dt, 1e3 * dt, M / (dev_ms * 1e-3), dev_ms, 2.0 * M * N * (1 if F == 1 else 2) / (dev_ms * 1e-3) / 1e12)
    if M <= 3_000_000:
        g = np.meshgrid(ax, ay, az, indexing="ij")
        Xs = np.ascontiguousarray(np.hstack([np.stack([gi.ravel() for gi in g], 1), np.full((M, 1), float(F - 1))]))
        core.predict(Xs, L.MEAN_ONLY)
        t0 = time.perf_counter(); m0, _ = core.predict(Xs, L.MEAN_ONLY); dt0 = time.perf_counter() - t0
        line += "; general mean-only %.3g pts/s; max|diff| %.2e" % (M / dt0, float(np.max(np.abs(m.ravel() - m0))))
    print(line, flush=True)
    core.close()
