"""Seeded inputs of the BASELINE-size parity cases (tests/test_gpu_scale_parity.py) and the oracle evaluation of each.
Shared by the tests and by ``oracle/make_golden_scale.py``, which freezes the oracle's outputs in
``tests/golden/scale_oracle.npz`` so that the GPU box does not spend minutes of host time re-deriving them (set
``GPC_SCALE_RECOMPUTE=1`` to run the oracle live instead).  Test infrastructure only."""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def sha(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a, dtype=np.float64).tobytes())
    return h.hexdigest()


# ---- configs[1]: 2-fidelity AR1, N = 2048, slice of the 100^3 grid + points next to training inputs -------------------
def configs1_inputs():
    import bench
    N, F = 2048, 2
    X4, y = bench.make_train(N, F)
    g = bench.make_grid(100, F - 1)
    Xs4 = np.vstack([np.ascontiguousarray(g[::50]), np.hstack([X4[:, :3] + 1e-3, np.ones((N, 1))]),
                     np.hstack([X4[:512, :3] - 1e-3, np.zeros((512, 1))])])
    return X4, y, bench.MF2_PARAMS.copy(), np.ascontiguousarray(Xs4)


def configs1_oracle():
    from oracle import gp_oracle as go
    X4, y, p, Xs4 = configs1_inputs()
    ref = go.MFGP(X4, y, p, F=2, gram=False)
    mu, var = ref.predict(Xs4)
    refg = go.MFGP(X4, y, p, F=2, gram=True)      # GPy's Gram-trick distances: the CPU-vs-CPU formulation spread
    mug, varg = refg.predict(Xs4)
    return dict(c1_sha=sha(X4, y, p, Xs4), c1_mu=mu[:, 0], c1_var=var[:, 0], c1_nlml=ref.f.nlml,
                c1_mu_gram=mug[:, 0], c1_var_gram=varg[:, 0])


# ---- configs[3]: N = 4096, F = 3, k = 32 information gain ---------------------------------------------------------------
def configs3_inputs():
    import bench
    N, F, k = 4096, 3, 32
    X4, y = bench.make_train(N, F, seed=3)
    rows, offs = bench.make_candidates(64, k, F)
    cands = [rows[offs[c]:offs[c + 1]] for c in (0, 9, 17, 23, 31, 42, 55, 63)]
    g = np.meshgrid(np.linspace(0, 10, 10), np.linspace(0, 20, 6), np.linspace(0, 10, 5))
    grid4 = np.ascontiguousarray(np.hstack([np.array([gi.ravel("F") for gi in g]).T, 2 * np.ones((300, 1))]))
    return X4, y, bench.MF3_PARAMS.copy(), cands, grid4


def configs3_oracle(n_seq_literal=2):
    """Literal refit loops: sequential operator on ``n_seq_literal`` candidates (32 refits of N + i rows each) and on all
    eight through the Schur form; log-det operators (plain, emukit-clipped) literally on all eight."""
    from oracle import gp_oracle as go
    X4, y, p, cands, grid4 = configs3_inputs()
    # MFGPCached: identical matrices, only their ASSEMBLY is incremental; every refit is a full factorisation
    ref = go.MFGPCached(X4, y, p, F=3, gram=False)
    sig_n = float(p[-1])
    _, _, _, noise = go.split_mf_params(p, 3)
    seq_schur = np.array([go.ig_seq_schur(ref, c, noise[c[:, 3].astype(int)], noise[0], sig_n,
                                          Xpred=np.hstack([c[:, :3], np.zeros((len(c), 1))])) for c in cands])
    seq_loop = np.array([go.ig_seq_mf_refit(ref, c, sig_n, 0) for c in cands[:n_seq_literal]])
    ld_loop = np.array([go.ig_logdet_refit(ref, grid4, c) for c in cands])
    ldc_loop = np.array([go.ig_logdet_refit(ref, grid4, c, clip_cov=1e-10) for c in cands])
    return dict(c3_sha=sha(X4, y, p, grid4, *cands), c3_seq_schur=seq_schur, c3_seq_loop=seq_loop, c3_ld_loop=ld_loop,
                c3_ldc_loop=ldc_loop)


# ---- N = 16384 single fidelity ----------------------------------------------------------------------------------------
def sf16384_inputs():
    import bench
    N, M = 16384, 1000
    X4, y = bench.make_train(N, 3, seed=16)
    X = np.ascontiguousarray(X4[:, :3])
    p = np.array([4.0, 2.0, 3.0, 2.5, 0.05])
    Xs = np.random.default_rng(99).uniform([0, 0, 0], [10, 20, 10], (M, 3))
    Xs[:200] = X[:200] + 1e-3
    return X, y, p, Xs


def sf16384_oracle():
    from oracle import gp_oracle as go
    X, y, p, Xs = sf16384_inputs()
    ref = go.SFGP(X, y, p, gram=False)
    mu, var = ref.predict(Xs)
    return dict(s16_sha=sha(X, y, p, Xs), s16_mu=mu[:, 0], s16_var=var[:, 0], s16_nlml=ref.f.nlml)


def frozen_or_live(prefix, live_fn, want_sha):
    """The oracle's outputs for one case: the committed fixture when it was generated on exactly these inputs, else a
    live evaluation (also with GPC_SCALE_RECOMPUTE=1)."""
    path = os.path.join(ROOT, "tests", "golden", "scale_oracle.npz")
    if not os.environ.get("GPC_SCALE_RECOMPUTE") and os.path.exists(path):
        g = np.load(path)
        if prefix + "_sha" in g and str(g[prefix + "_sha"]) == want_sha:
            return {k: g[k] for k in g.files if k.startswith(prefix + "_")}, "frozen"
    return live_fn(), "live"
