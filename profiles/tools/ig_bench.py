import sys, os, time, numpy as np
sys.path.insert(0, os.getcwd())
import __graft_entry__ as e; e.setup_path()
import gpcore, bench, torch
from gpcore import _lib as L
C = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
N, F, k = 4096, 3, 32
X4, y = bench.make_train(N, F, seed=3)
core = gpcore.GPCore(L.KIND_MF_AR1_RBF, F, 0)
core.set_hypers(bench.MF3_PARAMS, 1e-8); core.set_data(X4, y); core.factor()
g = np.meshgrid(np.linspace(0, 10, 10), np.linspace(0, 20, 6), np.linspace(0, 10, 5))
grid4 = np.ascontiguousarray(np.hstack([np.array([gi.ravel("F") for gi in g]).T, 2 * np.ones((300, 1))]))
rows, offs = bench.make_candidates(C, k, F)
core.ig_logdet(grid4, rows[:k * 1024], offs[:1025]); core.ig_logdet(grid4, rows[:k * 1024], offs[:1025], clip=True); core.ig_seq(rows[:k * 1024], offs[:1025], bench.MF3_PARAMS[-1], pred_fid=0)
for name, fn in (("logdet", lambda: core.ig_logdet(grid4, rows, offs)), ("seq", lambda: core.ig_seq(rows, offs, bench.MF3_PARAMS[-1], pred_fid=0)),
                 ("clip", lambda: core.ig_logdet(grid4, rows, offs, clip=True))):
    torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); dt = time.perf_counter() - t0
    print("%s: %.1f evals/s (%.1f ms for %d)" % (name, C / dt, 1e3 * dt, C), flush=True)
