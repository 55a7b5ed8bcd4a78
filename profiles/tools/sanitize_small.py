"""A small pass through every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool memcheck python profiles/tools/sanitize_small.py
factor (N = 300, 3 fidelities), posterior in all three modes, full covariance, information gain (all operators),
NLML gradients, evaluator, candidate generation is covered by the tests."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
entry.setup_path()
import gpcore
from gpcore import _lib as L

rng = np.random.default_rng(0)
N, M, F = 300, 700, 3
X4 = np.hstack([rng.uniform([0, 0, 0], [10, 20, 10], (N, 3)), rng.integers(0, F, (N, 1)).astype(float)])
y = np.sin(X4[:, 0]) + 0.3 * X4[:, 3] + 0.05 * rng.standard_normal(N)
p = np.array([3.0, 2.5, 3.5, 3.0, 1.0, 1.5, 2.0, 2.0, 0.5, 1.0, 1.5, 1.5, 0.9, 1.1, 0.08, 0.04, 0.02])
Xs4 = np.hstack([rng.uniform([0, 0, 0], [10, 20, 10], (M, 3)), 2.0 * np.ones((M, 1))])
core = gpcore.GPCore(L.KIND_MF_AR1_RBF, F, 0)
core.set_hypers(p, 1e-8); core.set_data(X4, y)
print("nlml", core.factor()[0])
core.factor()                                   # second call: CUDA-graph replay
flags = L.INCLUDE_NOISE | L.CLIP_DIAG
for mode in (L.MODE_INT8, L.MODE_FP64, L.MODE_INT8_F32):
    core.set_mode(mode)
    m, v = core.predict(Xs4, flags)
    print("mode", mode, float(m.sum()), float(v.sum()))
core.set_mode(L.MODE_INT8)
_, C = core.predict_cov(Xs4[:200], flags)
g = core.nlml_grad(p.size)
cands = [np.hstack([rng.uniform([0, 0, 0], [10, 20, 10], (k, 3)), rng.integers(0, F, (k, 1)).astype(float)]) for k in (3, 32, 0, 17, 64)]
rows, offs = gpcore.GPCore._ragged(cands)
grid4 = np.hstack([rng.uniform([0, 0, 0], [10, 20, 10], (60, 3)), 2.0 * np.ones((60, 1))])
print("seq", core.ig_seq(rows, offs, 0.02, pred_fid=0)[0])
print("logdet", core.ig_logdet(grid4, rows, offs)[0])
print("clip", core.ig_logdet(grid4, rows, offs, clip=True)[0])
print("self", core.ig_selfgrid(rows, offs, pred_fid=2)[0])
print("spd", core.spd_stats(C, rng.standard_normal(200)))
core.close()
print("done")
