import sys, os, time, numpy as np
sys.path.insert(0, os.getcwd())
import __graft_entry__ as e; e.setup_path()
import gpcore, bench
from gpcore import _lib as L
sizes = [int(a) for a in sys.argv[1:]] or [2048, 4096, 8192]
for N in sizes:
    X4, y = bench.make_train(N, 2)
    core = gpcore.GPCore(L.KIND_MF_AR1_RBF, 2, 0)
    core.set_hypers(bench.MF2_PARAMS, 1e-8); core.set_data(X4, y)
    core.factor()
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); nlml, ld = core.factor(); ts.append(time.perf_counter() - t0)
    t = min(ts)
    print("N=%d factor %.2f ms  (chol N^3/3 + inverse N^3/3 = %.2f TFLOP/s) nlml %.6f" % (N, 1e3 * t, 2 * N**3 / 3 / t / 1e12, nlml), flush=True)
    core.close()
