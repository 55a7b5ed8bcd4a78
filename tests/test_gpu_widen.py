"""GPU parity for the widened rows: calculatePathInfoEmu2 (self-grid log-det IG), the windowed
sequential SF operators (calcPathInfoSF / calcPathInfoSF4) against literal restatements of the
reference loops, and the trainer's evaluator (covariance-weighted MSE)."""
import numpy as np
import pytest

from conftest import golden, normwise

pytestmark = pytest.mark.gpu

TOL = 1e-9
MF_PARAMS = np.array([3.0, 2.5, 3.5, 3.0, 1.0, 1.5, 2.0, 2.0, 0.5, 1.0, 1.5, 1.5, 0.9, 1.1, 0.08, 0.04, 0.02])
SF_PARAMS = np.array([4.0, 2.0, 3.0, 2.5, 0.05])


@pytest.fixture(scope="module")
def gpcore_mod(built_lib):
    import gpcore
    return gpcore


@pytest.fixture(scope="module")
def go():
    from oracle import gp_oracle
    return gp_oracle


def test_ig_selfgrid_emu2(gpcore_mod, go):
    L_ = gpcore_mod._lib
    g = golden("gp_oracle.npz")
    rng = np.random.default_rng(41)
    core = gpcore_mod.GPCore(L_.KIND_MF_AR1_RBF, 3, 0)
    core.set_hypers(MF_PARAMS, 1e-8)
    core.set_data(g["X4"], g["y4"])
    core.factor()
    ref = go.MFGP(g["X4"], g["y4"], MF_PARAMS, F=3, gram=False)
    ks = [1, 4, 9, 16, 0, 30]
    cands = [np.hstack([rng.uniform([0, 0, 0], [10, 20, 10], (k, 3)), rng.integers(0, 3, (k, 1)).astype(float)]) for k in ks]
    cands.append(np.hstack([rng.uniform([0, 0, 0], [10, 20, 10], (6, 3)), 2 * np.ones((6, 1))]))   # all at pred_fid: no query copies
    rows, offs = gpcore_mod.GPCore._ragged(cands)
    for clip in (True, False):
        I, best = core.ig_selfgrid(rows, offs, pred_fid=2, clip=clip)
        want = np.array([go.ig_selfgrid_refit(ref, c, 2, 1e-10 if clip else None) if len(c) else 0.0 for c in cands])
        assert normwise(I, want, 1.0) < 1e-8, (clip, I, want)
        assert best == int(np.argmax(want))
    core.close()


def _agent(gpcore_mod, X, y):
    from gpcore.GPy.kern import RBF
    from gpcore.GPy.models import GPRegression
    from gpcore.infogain import InfoGainOperators

    class Agent(InfoGainOperators):
        fidLevs = [0.25, 2.25, 6.25]

        def pathToTrajPoints(self, V, E, path, dense=False, t_off=0, withVar=False):
            p = E[path]
            return p if withVar else p[:, :4]

    ag = Agent()
    ag.sfgp = GPRegression(X, y[:, None], RBF(3, ARD=True))
    ag.sfgp.param_array[:] = SF_PARAMS
    return ag


def _literal_sf(go, X, y, pnts, first_windowed):
    """calcPathInfoSF (GraceRIGV3.py:468-503) / calcPathInfoSF4 (Phys/GraceRIGV3.py:498-534), literally."""
    lx, ly = SF_PARAMS[1], SF_PARAMS[2]
    sig_n = SF_PARAMS[-1]
    win = lambda A: A[np.logical_and(A[:, 0] < 3 * lx, A[:, 1] < 3 * ly)]
    x0 = pnts[:1, :3]
    allX = np.concatenate((X, x0))
    if first_windowed:
        tempX = win(allX)
        gp = go.SFGP(np.concatenate((tempX, x0)), np.zeros(len(tempX) + 1), SF_PARAMS, gram=False)
    else:
        gp = go.SFGP(np.concatenate((X, x0)), np.concatenate((y, [0.0])), SF_PARAMS, gram=False)
    I = np.log(1 + gp.predict(x0)[1][0, 0] / sig_n)
    for i in range(1, len(pnts)):
        xi = pnts[i:i + 1, :3]
        allX = np.concatenate((allX, xi))
        tempX = allX.copy()
        if allX.shape[0] > 100:
            tempX = win(allX)
            tempX = allX if tempX.shape[0] == 0 else tempX
        gp = go.SFGP(tempX, np.zeros(len(tempX)), SF_PARAMS, gram=False)
        I += np.log(1 + gp.predict(xi)[1][0, 0] / sig_n)
    return I


@pytest.mark.parametrize("first_windowed", [False, True])
def test_sf_windowed_operators(gpcore_mod, go, first_windowed):
    d = golden("field_data.npz")
    X, y = d["Xh"][:300], d["y"][:300]
    ag = _agent(gpcore_mod, X, y)
    rng = np.random.default_rng(43)
    E = {}
    starts = [[1.0, 2.0, 3.0], [5.5, 8.5, 2.0], [8.0, 15.0, 5.0], [5.9, 8.9, 4.0]]   # inside, straddling, outside the window
    for c, a in enumerate(starts):
        kk = 5 + 2 * c
        pts = np.asarray(a)[None] + np.linspace(0, 1, kk)[:, None] * rng.normal(0, 1.5, 3)[None]
        E[c] = np.hstack([pts, np.arange(kk)[:, None]])
    E[9] = E[0][:1]
    paths = [0, 1, 2, 3, 9]
    op = "calcPathInfoSF4" if first_windowed else "calcPathInfoSF"
    I, best = ag.score_many(None, E, paths, operator=op)
    want = np.array([_literal_sf(go, X, y, E[c], first_windowed) for c in paths[:4]])
    assert normwise(I[:4], want) < 1e-8, (I, want)
    assert I[4] == -np.inf and best == int(np.argmax(want))
    assert abs(getattr(ag, op)(None, E, 1) - want[1]) < 1e-8 * abs(want[1])


def test_weighted_mse_evaluator(gpcore_mod, go):
    """GPTrainers.py:121-137 on a 2000-point grid covariance, against np.linalg.inv."""
    from gpcore import evaluate
    from gpcore.GPy.kern import RBF
    from gpcore.GPy.models import GPRegression
    d = golden("field_data.npz")
    gp = GPRegression(d["Xh"], d["y"][:, None], RBF(3, ARD=True))
    gp.param_array[:] = SF_PARAMS
    for pts in (d["test_sub"], d["test"]):
        mu, cov = gp.predict(pts, full_cov=1)
        err = mu - np.sin(pts[:, :1])
        got = evaluate.weighted_mse(err, cov)
        want = go.weighted_mse(err, cov)
        assert abs(got - want) < 1e-8 * abs(want), (got, want)
        got_raw = evaluate.weighted_mse(err, cov, normalize=False)
        want_raw = float((err.T @ np.linalg.inv(cov) @ err).item()) / len(err)
        assert abs(got_raw - want_raw) < 1e-8 * abs(want_raw)
    q, f, ld = gpcore_mod.GPCore(0, 1, 0).spd_stats(np.diag([1.0, 4.0, 0.25]), np.array([1.0, 2.0, 3.0]))
    assert abs(q - (1 + 1 + 36)) < 1e-12 and abs(f - np.sqrt(1 + 1 / 16 + 16)) < 1e-12 and abs(ld) < 1e-12
    with pytest.raises(np.linalg.LinAlgError):
        gpcore_mod.GPCore(0, 1, 0).spd_stats(np.array([[1.0, 2.0], [2.0, 1.0]]))
    assert abs(evaluate.rmse(np.array([3.0, 4.0])) - np.sqrt(12.5)) < 1e-15


def _fd_grad(f, p, rel=1e-6):
    g = np.zeros_like(p)
    for i in range(p.size):
        h = rel * max(abs(p[i]), 1e-3)
        a, b = p.copy(), p.copy()
        a[i] += h; b[i] -= h
        g[i] = (f(a) - f(b)) / (2 * h)
    return g


@pytest.mark.parametrize("case", ["sf_rbf", "sf_mat32", "mf3_mixed", "mf3_single", "mf2_mat32"])
def test_analytic_nlml_gradients(gpcore_mod, go, case):
    """gpc_nlml_grad against central differences of the ORACLE's NLML (CPU, float64)."""
    L_ = gpcore_mod._lib
    rng = np.random.default_rng(51)
    N = 260
    if case.startswith("sf"):
        kind, ok = (L_.KIND_SF_RBF, go.KIND_RBF) if case == "sf_rbf" else (L_.KIND_SF_MAT32, go.KIND_MAT32)
        X = rng.uniform([0, 0, 0], [10, 20, 10], (N, 3)); y = np.sin(X[:, 0]) + 0.1 * rng.standard_normal(N)
        p = np.array([2.0, 1.5, 2.5, 2.0, 0.07])
        X4 = np.hstack([X, np.zeros((N, 1))]); F = 1
        f = lambda q: go.SFGP(X, y, q, kind=ok, gram=False).f.nlml
    else:
        F = 2 if case == "mf2_mat32" else 3
        ok = go.KIND_MAT32 if case == "mf2_mat32" else go.KIND_RBF
        kind = L_.KIND_MF_AR1_MAT32 if case == "mf2_mat32" else L_.KIND_MF_AR1_RBF
        X4 = np.hstack([rng.uniform([0, 0, 0], [10, 20, 10], (N, 3)), rng.integers(0, F, (N, 1)).astype(float)])
        y = np.sin(X4[:, 0]) + 0.3 * X4[:, 3] + 0.1 * rng.standard_normal(N)
        if F == 3:
            p = np.array([3.0, 2.5, 3.5, 3.0, 1.0, 1.5, 2.0, 2.0, 0.5, 1.0, 1.5, 1.5, 0.9, 1.1, 0.08, 0.04, 0.02])
            if case == "mf3_single":
                p = p[:15]
        else:
            p = np.array([4.0, 2.0, 3.0, 2.5, 1.0, 1.5, 2.0, 2.0, 0.8, 0.05, 0.02])
        f = lambda q: go.MFGP(X4, y, q, F=F, kind=ok, gram=False).f.nlml
    core = gpcore_mod.GPCore(kind, F, 0)
    core.set_hypers(p, 1e-8)
    core.set_data(X4, y)
    core.factor()
    g, dW = core.nlml_grad(p.size, want_diag=True)
    want = _fd_grad(f, p)
    assert np.max(np.abs(g - want) / np.maximum(np.abs(want), 1.0)) < 2e-6, (g, want)
    Ky_inv_minus = np.linalg.inv((go.SFGP(X4[:, :3], y, p, kind=ok, gram=False).f.L if F == 1 else
                                  go.MFGP(X4, y, p, F=F, kind=ok, gram=False).f.L))
    ref = go.SFGP(X4[:, :3], y, p, kind=ok, gram=False).f if F == 1 else go.MFGP(X4, y, p, F=F, kind=ok, gram=False).f
    Wd = np.sum(Ky_inv_minus ** 2, axis=0) - ref.alpha ** 2
    assert normwise(dW, Wd) < 1e-8
    core.close()


def test_model_gradients_and_optimize(gpcore_mod, go):
    """Mirror models: objective_function_gradients in param_array order (shared lengthscale sums the
    per-dimension terms; fixed scale stays fixed), optimize() lowers the NLML, NIGP analytic gradient."""
    from gpcore.GPy.kern import RBF
    from gpcore.GPy.models import GPRegression
    from gpcore.emukit.multi_fidelity.kernels import LinearMultiFidelityKernel
    from gpcore.emukit.multi_fidelity.models import GPyLinearMultiFidelityModel
    from gpcore.emukit.model_wrappers.gpy_model_wrappers import GPyMultiOutputWrapper
    from gpcore import nigp
    d, g = golden("field_data.npz"), golden("gp_oracle.npz")
    X, y = d["Xh"][:300], d["y"][:300]
    m = GPRegression(X, y[:, None], RBF(3, variance=2.0, lengthscale=1.7))     # one shared lengthscale
    m.Gaussian_noise.variance = 0.1
    gm = m.objective_function_gradients()
    assert gm.shape == (3,)
    f = lambda q: go.SFGP(X, y, np.array([q[0], q[1], q[1], q[1], q[2]]), gram=False).f.nlml
    want = _fd_grad(f, m.param_array.copy())
    assert np.max(np.abs(gm - want) / np.maximum(np.abs(want), 1.0)) < 2e-6
    m2 = GPRegression(X, y[:, None], RBF(3, ARD=True))
    f0 = m2.objective_function()
    m2.optimize(max_iters=40)
    assert m2.objective_function() < f0 - 1.0 and np.all(m2.param_array > 0)
    assert np.max(np.abs(m2.objective_function_gradients())) < 5.0
    # multi-fidelity with the scale fixed (GPTrainers.py:67) -- it must not move
    sel = np.r_[0:80, 400:480, 600:680]
    k = LinearMultiFidelityKernel([RBF(3, ARD=True) for _ in range(3)])
    w = GPyMultiOutputWrapper(GPyLinearMultiFidelityModel(g["X4"][sel], g["y4"][sel, None], k, n_fidelities=3), 3, 1)
    w.gpy_model.kern.scale.fix([1, 1])
    f0 = w.gpy_model.objective_function()
    w.optimize()
    assert w.gpy_model.objective_function() < f0 - 1.0
    assert list(w.gpy_model.kern.scale) == [1.0, 1.0]
    # NIGP: analytic gradient w.r.t. log hypers vs central differences of the module objective
    gd = golden("nigp_demo.npz")
    lh = gd["log_hyp"] + 0.05
    v, gr = nigp.nlml_and_grad(lh, gd["X"], gd["y"], gd["grads"])
    assert abs(v - nigp.neg_log_marginal_likelihood(lh, gd["X"], gd["y"], gd["grads"])) < 1e-9
    fd = _fd_grad(lambda q: go.nigp_nlml(q, gd["X"], gd["y"], gd["grads"], gram=False), lh, rel=1e-5)
    assert np.max(np.abs(gr - fd) / np.maximum(np.abs(fd), 1.0)) < 1e-5, (gr, fd)
    np.random.seed(1)
    a = nigp.NIGP(n_restarts=1, iters=2, verbose=False).fit(gd["X"], gd["y"], maxiter_opt=50)
    assert np.all(np.isfinite(a.get_params())) and a.predict(gd["Xs"])[1].min() >= 1e-12


def test_config1_gptrainers_flow_matches_published_results(gpcore_mod):
    """BASELINE configs[0]: the reference's GPTrainers.py flow (fit MF / SF / SF-true-position / NIGP,
    predict on the 2000-point grid, RMSE + covariance-weighted MSE) on the bundled dataset
    GPData_0.2_fieldMeas_0_T0_0, against the numbers the reference PUBLISHED for it
    (Data/TrajectoriesAndEstimates/GPResults/MSE_0.2_fieldMeas_0_T0_0.txt).  Those depend on the
    optimiser's path, so this is a band (1e-4 relative on RMSE), not a 1e-9 pin."""
    import importlib.util
    import os
    from conftest import ROOT
    spec = importlib.util.spec_from_file_location("gptrainers_flow", os.path.join(ROOT, "examples", "gptrainers_flow.py"))
    flow = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(flow)
    d = golden("field_data.npz")
    keep = d["t"] < 3600
    cols = {"t": d["t"][keep], "x": d["X"][keep, 0], "y": d["X"][keep, 1], "z": d["X"][keep, 2],
            "xh": d["Xh"][keep, 0], "yh": d["Xh"][keep, 1], "zh": d["Xh"][keep, 2],
            "fieldVal": d["y"][keep], "fidLev": d["fidLev"][keep]}
    np.random.seed(0)
    rm, wm = flow.run(cols, verbose=False)
    published = {"mf": 5.248310997830455, "sf": 5.247514252892991, "nisf": 5.247403702851988, "sfTP": 5.243211870740189}
    for k, v in published.items():
        assert abs(rm[k] - v) < 1e-4 * v, (k, rm[k], v)
    assert abs(wm["sf"] - 0.07326736) < 1e-3 * 0.07326736
    assert abs(wm["sfTP"] - 0.07317822) < 1e-3 * 0.07317822
    assert wm["nisf"] < 1e-6


def _random_paths(rng, n_paths):
    paths = []
    for _ in range(n_paths):
        ne = int(rng.integers(1, 4))
        pos = rng.uniform([0, 0], [10, 20])
        path = []
        for _ in range(ne):
            nxt = pos + rng.normal(0, 2.0, 2)
            prims, depth = [], 0.0
            for _ in range(int(rng.integers(1, 6))):
                kind = int(rng.integers(0, 4))
                if kind == 0:
                    dz = float(rng.uniform(-1, 2)); prims.append((0.0, dz, 0.0, float(rng.uniform(0.05, 0.3))))
                elif kind == 1:
                    dz = float(rng.uniform(-1.5, 2)); prims.append((1.0, float(rng.uniform(0.3, 1.0)), dz, float(rng.uniform(0.05, 0.3))))
                elif kind == 2:
                    prims.append((2.0, float(rng.uniform(0.2, 3.0)), float(rng.uniform(0.1, 0.5)), 0.0))
                else:
                    dz = float(rng.uniform(-1, 2)); prims.append((3.0, dz, float(rng.uniform(0.05, 0.3)), 0.0))
            path.append((pos.copy(), nxt.copy(), prims))
            pos = nxt
        paths.append(path)
    return paths


@pytest.mark.parametrize("dense", [False, True])
@pytest.mark.parametrize("with_var", [True, False])
def test_candidate_generation_matches_reference_restatement(gpcore_mod, dense, with_var):
    """Device evaluateTraj / edgePointsToTrajPoints / pathToTrajPoints + fidelity labels against the
    NumPy restatement of GraceRIGV3.py:235-294,373-427 (oracle/traj_oracle.py)."""
    from gpcore import trajectory
    from gpcore.infogain import label_fidelity
    from oracle import traj_oracle as to
    rng = np.random.default_rng(61)
    paths = _random_paths(rng, 40)
    paths.append([((1.0, 1.0), (2.0, 2.0), [(2.0, 0.0, 0.3, 0.0)])])          # zero-length swim: duplicate way-points
    paths.append([((1.0, 1.0), (1.0, 3.0), [(3.0, 1.0, 0.1, 0.0), (3.0, -1.0, 0.1, 0.0)]),
                  ((1.0, 3.0), (1.0, 1.0), [(3.0, 1.0, 0.1, 0.0), (3.0, -1.0, 0.1, 0.0)])])   # there and back again
    fl = [0.05, 0.15, 0.3]
    pts, fids = trajectory.paths_to_points(paths, 0.01, 0.2, dense=dense, with_var=with_var, fid_levels=fl, max_pts=256)
    for c, path in enumerate(paths):
        want = to.path_to_traj_points(path, 0.01, 0.2, dense=dense, with_var=with_var)
        got = pts[c][:, :want.shape[1]]
        assert got.shape == want.shape, (c, got.shape, want.shape)
        assert np.max(np.abs(got - want), initial=0.0) < 1e-11, c
        if with_var:
            assert np.array_equal(fids[c], label_fidelity(want[:, 4], fl))
    rows = trajectory.candidate_rows(paths[:5], 0.01, 0.2, fl, dense=dense, max_pts=256)
    assert rows[0].shape[1] == 4 and set(np.unique(np.concatenate(rows)[:, 3])) <= {0.0, 1.0, 2.0}
    with pytest.raises(ValueError):
        trajectory.paths_to_points(paths[:3], 0.01, 5.0, dense=True, max_pts=4)
    assert trajectory.paths_to_points([], 0.01, 0.2)[0] == []


@pytest.mark.parametrize("dense", [0, 1])
@pytest.mark.parametrize("wv", [0, 1])
def test_candidate_generation_matches_reference_planner_golden(gpcore_mod, dense, wv):
    """Device candidate generation against outputs of the reference's OWN GraceAgent.pathToTrajPoints
    (tests/golden/traj_paths.npz -- pinned)."""
    from gpcore import trajectory
    g = golden("traj_paths.npz")
    eo, po = g["edge_off"], g["prim_off"]
    paths = [[(g["edge_xy"][e][:2], g["edge_xy"][e][2:], [tuple(p) for p in g["prims"][po[e]:po[e + 1]]])
              for e in range(eo[c], eo[c + 1])] for c in range(len(eo) - 1)]
    pts, _ = trajectory.paths_to_points(paths, float(g["variance_rate"]), float(g["meas_rate"]), dense=bool(dense),
                                        with_var=bool(wv), max_pts=256)
    want, off = g["pts_d%d_v%d" % (dense, wv)], g["off_d%d_v%d" % (dense, wv)]
    for c in range(len(paths)):
        w = want[off[c]:off[c + 1]]
        got = pts[c][:, :w.shape[1]]
        assert got.shape == w.shape and np.max(np.abs(got - w), initial=0.0) < 1e-11, c


def test_info_gain_operators_match_reference_operator_code(gpcore_mod):
    """Full device pipeline (paths -> candidate points -> information gain) against golden values produced
    by the reference's OWN operators (root and PhysicalExperimentCode GraceRIGV3.py run unmodified over
    adapters of the restated GP models; tests/golden/ig_operators.npz)."""
    from gpcore import trajectory
    from gpcore.GPy.kern import RBF
    from gpcore.GPy.models import GPRegression
    from gpcore.emukit.multi_fidelity.kernels import LinearMultiFidelityKernel
    from gpcore.emukit.multi_fidelity.models import GPyLinearMultiFidelityModel
    from gpcore.emukit.model_wrappers.gpy_model_wrappers import GPyMultiOutputWrapper
    from gpcore.infogain import InfoGainOperators
    g, t = golden("ig_operators.npz"), golden("traj_paths.npz")
    eo, po = t["edge_off"], t["prim_off"]
    edges = {c: [(t["edge_xy"][e][:2], t["edge_xy"][e][2:], [tuple(p) for p in t["prims"][po[e]:po[e + 1]]])
                 for e in range(eo[c], eo[c + 1])] for c in range(int(g["n_paths"]))}

    class Agent(InfoGainOperators):
        fidLevs = list(g["fidLevs"])

        def pathToTrajPoints(self, V, E, path, dense=False, t_off=0, withVar=False):
            pts, _ = trajectory.paths_to_points([E[path]], float(g["variance_rate"]), float(g["meas_rate"]), dense=dense,
                                                with_var=withVar, t_off=t_off, max_pts=256)
            return pts[0] if withVar else pts[0][:, :4]

    ag = Agent()
    ag.fieldGrid = g["grid"]
    ag.sfgp = GPRegression(g["Xh"], g["y"][:, None], RBF(3, ARD=True))
    ag.sfgp.param_array[:] = g["sf_params"]
    k = LinearMultiFidelityKernel([RBF(3, ARD=True) for _ in range(3)])
    ag.mfgp = GPyMultiOutputWrapper(GPyLinearMultiFidelityModel(g["X4"], g["y"][:, None], k, n_fidelities=3), 3, 1)
    ag.mfgp.gpy_model.param_array[:] = g["mf_params"]
    paths = list(edges)
    cases = [("root_calcPathInfoSF2", "calcPathInfoSF2", {}),
             ("root_calcPathInfoSF", "calcPathInfoSF", {}),
             ("phys_calcPathInfoSF4", "calcPathInfoSF4", {}),
             ("root_calculatePathInfoEmu", "calculatePathInfoEmu", {"sig_index": -3}),
             ("phys_calculatePathInfoEmu", "calculatePathInfoEmu", {"sig_index": -1}),
             ("root_calculatePathInfoEmu2", "calculatePathInfoEmu2", {}),
             ("phys_calcPathInfoSFBatch", "calcPathInfoSFBatch", {}),
             ("phys_calculatePathInfoEmuBatch", "calculatePathInfoEmuBatch", {})]
    for gold, op, kw in cases:
        ag.logDetPrior = None
        I, best = ag.score_many(None, edges, paths, operator=op, **kw)
        want = g[gold]
        assert normwise(I, want, 1.0) < 1e-7, (gold, I, want)
        assert best == int(np.argmax(want)), gold
    # the single-path calcPathInfoSFBatch as the planner drives it: consecutive calls after one reset keep the
    # points of the earlier calls in the cached copy (reference quirk, reproduced by default)
    ag.logDetPrior = None
    ag._sfb_model = None
    got = np.array([ag.calcPathInfoSFBatch(None, edges, c) for c in paths])
    assert normwise(got, g["phys_calcPathInfoSFBatch_consecutive"], 1.0) < 1e-7, (got, g["phys_calcPathInfoSFBatch_consecutive"])
    assert np.max(np.abs(got[1:] - g["phys_calcPathInfoSFBatch"][1:])) > 1e-3      # and differs from independent scoring
    # small training sets: the <= 100-row branches of the windowed operators (the switch to the window falls
    # inside a path at N = 60 and N = 97) and a training set with no row inside the window
    small = {"root_calcPathInfoSF": ("calcPathInfoSF", {}), "phys_calcPathInfoSF4": ("calcPathInfoSF4", {}),
             "root_calculatePathInfoEmu": ("calculatePathInfoEmu", {"sig_index": -3}),
             "phys_calculatePathInfoEmu": ("calculatePathInfoEmu", {"sig_index": -1})}
    for tag in ("n60", "n97", "nowin", "nowin40"):
        idx = g[tag + "_idx"]
        ag.sfgp.set_XY(g["Xh"][idx], g["y"][idx][:, None])
        ag.mfgp.set_data(g["X4"][idx], g["y"][idx][:, None])
        for gold, (op, kw) in small.items():
            if tag + "_" + gold not in g:
                continue
            I, best = ag.score_many(None, edges, paths, operator=op, **kw)
            want = g[tag + "_" + gold]
            assert normwise(I, want, 1.0) < 1e-7, (tag, gold, I, want)
            assert best == int(np.argmax(want)), (tag, gold)


def test_ig_logdet_emukit_clip(gpcore_mod, go):
    """calculatePathInfoEmuBatch takes both determinants of emukit's predict_covariance, i.e. of G x G
    matrices clipped element-wise at 1e-10 (PhysicalExperimentCode/GraceRIGV3.py:608-617).  The device path
    with GPC_CLIP_COV against the literal refit loop of the oracle (clip active: the posterior covariance
    between distant grid points is negative), on a small grid and on the reference's 10 x 6 x 5 grid."""
    L_ = gpcore_mod._lib
    g = golden("gp_oracle.npz")
    rng = np.random.default_rng(77)
    core = gpcore_mod.GPCore(L_.KIND_MF_AR1_RBF, 3, 0)
    core.set_hypers(MF_PARAMS, 1e-8)
    core.set_data(g["X4"], g["y4"])
    core.factor()
    ref = go.MFGP(g["X4"], g["y4"], MF_PARAMS, F=3, gram=False)
    ks = [1, 5, 0, 17, 32, 33, 64]
    cands = [np.hstack([rng.uniform([0, 0, 0], [10, 20, 10], (k, 3)), rng.integers(0, 3, (k, 1)).astype(float)]) for k in ks]
    rows, offs = gpcore_mod.GPCore._ragged(cands)
    ax = [np.linspace(0, 10, 10), np.linspace(0, 20, 6), np.linspace(0, 10, 5)]
    big = np.stack(np.meshgrid(*ax, indexing="ij"), -1).reshape(-1, 3)
    for grid in (rng.uniform([0, 0, 0], [10, 20, 10], (45, 3)), big):
        g4 = np.hstack([grid, 2 * np.ones((len(grid), 1))])
        cov0 = ref.predict_covariance(g4, clip_cov=None)
        assert (cov0 < 1e-10).any()                         # the clip is active on this problem
        I, prior, best = core.ig_logdet(g4, rows, offs, clip=True)
        want = np.array([go.ig_logdet_refit(ref, g4, c, clip_cov=1e-10) if len(c) else 0.0 for c in cands])
        assert abs(prior - np.linalg.slogdet(ref.predict_covariance(g4, clip_cov=1e-10))[1]) < 1e-8 * abs(prior)
        assert normwise(I, want, 1.0) < 1e-8, (I, want)
        assert best == int(np.argmax(want))
        J, _, _ = core.ig_logdet(g4, rows, offs, clip=False)  # and the unclipped value differs
        assert np.max(np.abs(J - I)) > 1e-6
    core.close()


def test_getEID_on_device_models(gpcore_mod):
    """The EID map through the mirrored SF / MF models (device predict) against the reference's getEID golden."""
    from gpcore.eid import getEID
    from gpcore.GPy.kern import RBF
    from gpcore.GPy.models import GPRegression
    from gpcore.emukit.multi_fidelity.kernels import LinearMultiFidelityKernel
    from gpcore.emukit.multi_fidelity.models import GPyLinearMultiFidelityModel
    from gpcore.emukit.model_wrappers.gpy_model_wrappers import GPyMultiOutputWrapper
    g = golden("eid.npz")
    sf = GPRegression(g["Xh"], g["y"][:, None], RBF(3, ARD=True))
    sf.param_array[:] = g["sf_params"]
    k = LinearMultiFidelityKernel([RBF(3, ARD=True) for _ in range(3)])
    mf = GPyMultiOutputWrapper(GPyLinearMultiFidelityModel(g["X4"], g["y"][:, None], k, n_fidelities=3), 3, 1)
    mf.gpy_model.param_array[:] = g["mf_params"]
    for auto in (0, 1):
        E, grid = getEID(sf, g["WS"], float(g["mD"]), auto=auto)
        assert np.max(np.abs(E - g["sf_auto%d" % auto])) < 1e-8 * np.max(g["sf_auto%d" % auto])
        E, _ = getEID(mf, g["WS"], float(g["mD"]), emu=True, auto=auto)
        assert np.max(np.abs(E - g["mf_auto%d" % auto])) < 1e-8 * np.max(g["mf_auto%d" % auto])


def test_tensor_grid_mean_as_gemm(gpcore_mod, go):
    """gpc_predict_grid_mean (separable squared-exponential cross-covariance, contraction over the training index as
    FP64 tensor-core GEMMs) against the general predict on the materialised grid and against the oracle: single
    fidelity, three-fidelity AR1 at every query fidelity, ragged axis lengths, several passes (> 8192 (ix, iy)
    pairs); Matern is rejected."""
    L_ = gpcore_mod._lib
    rng = np.random.default_rng(9)

    def mesh(ax, ay, az, f):
        g = np.meshgrid(ax, ay, az, indexing="ij")
        P = np.stack([gi.ravel() for gi in g], 1)
        return np.ascontiguousarray(np.hstack([P, np.full((len(P), 1), float(f))]))

    for kind, F, p, N in ((L_.KIND_SF_RBF, 1, SF_PARAMS, 333), (L_.KIND_MF_AR1_RBF, 3, MF_PARAMS, 700)):
        X4 = np.hstack([rng.uniform([0, 0, 0], [10, 20, 10], (N, 3)), rng.integers(0, F, (N, 1)).astype(float)])
        y = np.sin(X4[:, 0]) + 0.3 * X4[:, 3] + 0.05 * rng.standard_normal(N)
        core = gpcore_mod.GPCore(kind, F, 0)
        core.set_hypers(p, 1e-8)
        core.set_data(X4, y)
        core.factor()
        ref = go.SFGP(X4[:, :3], y, p, gram=False) if F == 1 else go.MFGP(X4, y, p, F=F, gram=False)
        shapes = [(7, 5, 3), (1, 1, 1), (33, 65, 70), (130, 90, 5)]          # the last one: 11700 pairs -> two passes
        for (nx, ny, nz) in shapes:
            ax, ay, az = np.sort(rng.uniform(0, 10, nx)), np.sort(rng.uniform(0, 20, ny)), np.linspace(0.5, 9.5, nz)
            for f in range(F):
                got = core.predict_grid_mean(ax, ay, az, fid=f)
                Xs = mesh(ax, ay, az, f)
                want, _ = core.predict(Xs, L_.MEAN_ONLY)
                assert got.shape == (nx, ny, nz)
                assert np.max(np.abs(got.ravel() - want)) <= 1e-11 * max(1.0, np.max(np.abs(want))), (kind, nx, ny, nz, f)
                if nx * ny * nz <= 200:
                    mu, _ = ref.predict(Xs[:, :3] if F == 1 else Xs)
                    assert np.max(np.abs(got.ravel() - mu[:, 0])) <= 1e-9 * max(1.0, np.max(np.abs(mu)))
        core.close()
    core = gpcore_mod.GPCore(L_.KIND_SF_MAT32, 1, 0)
    core.set_hypers(SF_PARAMS, 1e-8)
    core.set_data(X4, y)
    core.factor()
    with pytest.raises(Exception):
        core.predict_grid_mean([0.0, 1.0], [0.0], [0.0])
    core.close()
    # the mirrored models expose it next to predict(): emukit wrapper (query fidelity 2) and NIGP (return_var=False)
    from gpcore.GPy.kern import RBF
    from gpcore.emukit.multi_fidelity.kernels import LinearMultiFidelityKernel
    from gpcore.emukit.multi_fidelity.models import GPyLinearMultiFidelityModel
    from gpcore.emukit.model_wrappers.gpy_model_wrappers import GPyMultiOutputWrapper
    from gpcore.nigp import NIGP
    k = LinearMultiFidelityKernel([RBF(3, ARD=True) for _ in range(3)])
    mf = GPyMultiOutputWrapper(GPyLinearMultiFidelityModel(X4, y[:, None], k, n_fidelities=3), 3, 1)
    mf.gpy_model.param_array[:] = MF_PARAMS
    ax, ay, az = np.linspace(0, 10, 9), np.linspace(0, 20, 11), np.linspace(0, 10, 4)
    mu, _ = mf.predict(mesh(ax, ay, az, 2))
    assert np.max(np.abs(mf.predict_grid_mean(ax, ay, az, fid=2).ravel() - mu[:, 0])) < 1e-11 * max(1.0, np.max(np.abs(mu)))
    g, d = golden("nigp_field.npz"), golden("field_data.npz")
    m = NIGP(verbose=False)
    m.lengthscales_, m.sigma_f_, m.sigma_y_, m.sigma_x_ = g["ls"], float(g["sigma_f"]), float(g["sigma_y"]), g["sigma_x"]
    m.X_train_, m.y_train_, m.noise_diag_train_ = d["Xh"], d["y"], g["noise_diag"]
    want = m.predict(mesh(ax, ay, az, 0)[:, :3], return_var=False)
    assert np.max(np.abs(m.predict_grid_mean(ax, ay, az).ravel() - want)) < 1e-11 * max(1.0, np.max(np.abs(want)))


def test_reference_information_gain_script(gpcore_mod, go):
    """examples/information_gain_test.py = the reference's informationGainTest.py with the GPy import swapped; the
    same flow on the oracle's GPRegression restatement gives the numbers to compare with (1-D inputs, non-ARD RBF,
    assignment to Gaussian_noise.variance, set_XY to a single far-away prior point, full_cov predictions)."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("information_gain_test", os.path.join(root, "examples", "information_gain_test.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    for seed, n in ((0, 3), (1, 6), (2, 12)):
        I, I2, I3 = mod.run(seed, n, verbose=False)
        np.random.seed(seed)
        Xpred = np.array([np.arange(-3, 3, .1)]).T
        X = np.random.uniform(-3., 3., (n, 1))
        Y = np.sin(X) + np.random.randn(n, 1) * 0.05
        p = np.array([0.7407926234918235, 1.5704374366230516, 0.0010413149736387451])
        pad = lambda A: np.hstack([A, np.zeros((len(A), 2))])
        p5 = np.array([p[0], p[1], p[1], p[1], p[2]])
        ld = lambda K: np.linalg.slogdet(K)[1]
        g = go.SFGP(pad(np.array([[-100.0]])), np.array([[0.0]]), p5, gram=False)
        lp, lp2 = ld(g.predict(pad(Xpred), full_cov=True)[1]), ld(g.predict(pad(X), full_cov=True)[1])
        g.set_XY(pad(X), Y)
        wI = 0.5 * (lp - ld(g.predict(pad(Xpred), full_cov=True)[1]))
        wI3 = 0.5 * (lp2 - ld(g.predict(pad(X), full_cov=True)[1]))
        g.set_XY(pad(X[:1]), np.array([[0.0]]))
        wI2 = 0.5 * np.log(1 + g.predict(pad(X[:1]))[1][0, 0] / p[2])
        Xa = X[:1]
        for i in range(2, n):
            wI2 += 0.5 * np.log(1 + g.predict(pad(X[i:i + 1]))[1][0, 0] / p[2])
            Xa = np.concatenate((Xa, X[i:i + 1]))
            g.set_XY(pad(Xa), np.zeros((len(Xa), 1)))
        # the 60-point grid covariance is numerically singular (det ~ 1e-200): log-dets agree to ~1e-6 between LAPACKs
        assert abs(I - wI) < 1e-5 * max(1.0, abs(wI)), (seed, I, wI)
        assert abs(I2 - wI2) < 1e-8 * max(1.0, abs(wI2)), (seed, I2, wI2)
        assert abs(I3 - wI3) < 1e-7 * max(1.0, abs(wI3)), (seed, I3, wI3)


def test_optimize_with_an_active_bound(gpcore_mod):
    """``lengthscale.constrain_bounded(lo, hi)`` (``...MFGP.py:408-410,664-666``) with the bound ACTIVE: the optimum of
    the free problem lies outside, the bounded fit must end strictly inside the interval, lower the NLML, and leave a
    small projected gradient in the raw (Logistic) space -- not stall on a flat clamped objective."""
    from gpcore.GPy.kern import RBF
    from gpcore.GPy.models import GPRegression
    d = golden("field_data.npz")
    X, y = d["Xh"][:300], d["y"][:300]
    free = GPRegression(X, y[:, None], RBF(3, ARD=True))
    free.optimize(max_iters=60)
    ls_free = np.array(free.kern.lengthscale)
    hi = float(0.6 * ls_free.min())                          # every free optimum lies above the upper bound
    m = GPRegression(X, y[:, None], RBF(3, ARD=True))
    m.kern.lengthscale.constrain_bounded(1e-4, hi)
    f0 = m.objective_function()
    m.optimize(max_iters=60)
    ls = np.array(m.kern.lengthscale)
    assert np.all(ls > 1e-4) and np.all(ls < hi), (ls, hi)
    assert np.any(ls > 0.95 * hi)                             # pressed against the bound
    assert m.objective_function() < f0
    # finite-difference optimiser path agrees on where it lands (same transform, no analytic gradient)
    m2 = GPRegression(X, y[:, None], RBF(3, ARD=True))
    m2.kern.lengthscale.constrain_bounded(1e-4, hi)
    m2.optimize(max_iters=60, analytic_gradients=False)
    assert abs(m2.objective_function() - m.objective_function()) < 1e-2 * abs(m.objective_function())


def test_out_of_range_fidelity_is_rejected_everywhere(gpcore_mod):
    """emukit raises on a fidelity label outside [0, F); so must every entry point that takes fidelity-indexed rows
    (test rows of predict_cov, kernel_matrix rows and the information-gain grid were unchecked)."""
    L_ = gpcore_mod._lib
    rng = np.random.default_rng(3)
    N, F = 200, 3
    X4 = np.hstack([rng.uniform(0, 10, (N, 3)), rng.integers(0, F, (N, 1)).astype(float)])
    y = rng.standard_normal(N)
    p = np.array([3.0, 2.5, 3.5, 3.0, 1.0, 1.5, 2.0, 2.0, 0.5, 1.0, 1.5, 1.5, 0.9, 1.1, 0.08, 0.04, 0.02])
    core = gpcore_mod.GPCore(L_.KIND_MF_AR1_RBF, F, 0)
    core.set_hypers(p, 1e-8)
    core.set_data(X4, y)
    core.factor()
    good = np.hstack([rng.uniform(0, 10, (20, 3)), 2 * np.ones((20, 1))])
    cand, offs = good[:8].copy(), np.array([0, 8])
    for badval in (3.0, -1.0, 1.5, np.nan):
        bad = good.copy()
        bad[7, 3] = badval
        with pytest.raises(gpcore_mod.GpcoreError, match="fidelity"):
            core.predict(bad, L_.INCLUDE_NOISE)
        with pytest.raises(gpcore_mod.GpcoreError, match="fidelity"):
            core.predict_cov(bad, L_.INCLUDE_NOISE)
        with pytest.raises(gpcore_mod.GpcoreError, match="fidelity"):
            core.kernel_matrix(bad)
        with pytest.raises(gpcore_mod.GpcoreError, match="fidelity"):
            core.kernel_matrix(good, bad)
        with pytest.raises(gpcore_mod.GpcoreError, match="fidelity"):
            core.ig_logdet(bad, cand, offs)
        with pytest.raises(gpcore_mod.GpcoreError, match="fidelity"):
            core.ig_logdet(good, bad[:8], offs)
    # the handle stays usable, and the mirror model raises ValueError before reaching the device
    m, v = core.predict(good, L_.INCLUDE_NOISE)
    assert np.all(np.isfinite(m)) and np.all(v > 0)
    core.close()
    from gpcore.GPy.kern import RBF
    from gpcore.emukit.multi_fidelity.kernels import LinearMultiFidelityKernel
    from gpcore.emukit.multi_fidelity.models import GPyLinearMultiFidelityModel
    from gpcore.emukit.model_wrappers.gpy_model_wrappers import GPyMultiOutputWrapper
    k = LinearMultiFidelityKernel([RBF(3, ARD=True) for _ in range(F)])
    w = GPyMultiOutputWrapper(GPyLinearMultiFidelityModel(X4, y[:, None], k, n_fidelities=F), F, 1)
    bad = good.copy()
    bad[0, 3] = 3.0
    for call in (w.predict, w.predict_covariance, k.K):
        with pytest.raises(ValueError, match="fidelity"):
            call(bad)


def _inv_sym3(A):
    """Cofactor inverse of a 3 x 3 matrix, plain Python floats (no LAPACK, no oracle)."""
    (a, b, c), (d, e, f), (g, h, i) = A
    det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g)
    return [[(e * i - f * h) / det, (c * h - b * i) / det, (b * f - c * e) / det],
            [(f * g - d * i) / det, (a * i - c * g) / det, (c * d - a * f) / det],
            [(d * h - e * g) / det, (b * g - a * h) / det, (a * e - b * d) / det]], det


def test_ar1_closed_form_known_answers(gpcore_mod):
    """Known-answer test of the AR1 (Kennedy-O'Hagan) multi-fidelity posterior that does NOT go through
    oracle/gp_oracle.py: one training point per fidelity (F = 3), every quantity written out by hand from
        K[(x,i),(x',j)] = sum_{m <= min(i,j)} (prod_{l=m}^{i-1} rho_l)(prod_{l=m}^{j-1} rho_l) k_m(x, x'),
        k_m(x, x') = v_m exp(-1/2 sum_d ((x_d - x'_d) / l_md)^2),
    per-fidelity noise and GPy's 1e-8 jitter on the training diagonal; the 3 x 3 system is inverted by cofactors in
    plain Python.  Checks kern.K, predict (mean, noise-inclusive variance at each fidelity), predict_covariance
    (with emukit's 1e-10 clip) and the NLML."""
    import math
    from gpcore.GPy.kern import RBF
    from gpcore.emukit.multi_fidelity.kernels import LinearMultiFidelityKernel
    from gpcore.emukit.multi_fidelity.models import GPyLinearMultiFidelityModel
    from gpcore.emukit.model_wrappers.gpy_model_wrappers import GPyMultiOutputWrapper
    v = [2.0, 0.7, 0.3]
    ls = [[1.5, 2.0, 1.0], [0.8, 1.1, 0.9], [2.5, 0.6, 1.7]]
    rho = [0.8, 1.3]
    noise = [0.05, 0.02, 0.01]
    X = [[0.3, 1.0, 0.5], [0.9, 0.4, 0.2], [0.1, 0.7, 1.1]]       # training point of fidelity 0, 1, 2
    y = [0.7, -0.4, 1.2]
    xs = [[0.5, 0.8, 0.6], [0.2, 0.5, 0.9]]                        # two test locations

    def kb(m, a, b):
        return v[m] * math.exp(-0.5 * sum(((a[d] - b[d]) / ls[m][d]) ** 2 for d in range(3)))

    def coef(i, m):
        c = 1.0
        for l in range(m, i):
            c *= rho[l]
        return c

    def k(a, i, b, j):
        return sum(coef(i, m) * coef(j, m) * kb(m, a, b) for m in range(min(i, j) + 1))

    Ky = [[k(X[i], i, X[j], j) + ((noise[i] + 1e-8) if i == j else 0.0) for j in range(3)] for i in range(3)]
    Kinv, det = _inv_sym3(Ky)
    alpha = [sum(Kinv[i][j] * y[j] for j in range(3)) for i in range(3)]
    nlml = 0.5 * sum(y[i] * alpha[i] for i in range(3)) + 0.5 * math.log(det) + 1.5 * math.log(2 * math.pi)

    kern = LinearMultiFidelityKernel([RBF(3, ARD=True) for m in range(3)])
    X4 = np.array([X[i] + [float(i)] for i in range(3)])
    model = GPyLinearMultiFidelityModel(X4, np.array(y)[:, None], kern, n_fidelities=3)
    # param_array order pinned by PhysicalExperimentCode/...MFGP.py:670: (var, l(3)) x 3, rho1, rho2, noise x 3
    model.param_array[:] = np.concatenate([[v[0]], ls[0], [v[1]], ls[1], [v[2]], ls[2], rho, noise])
    w = GPyMultiOutputWrapper(model, 3, 1)
    Kdev = kern.K(X4)
    for i in range(3):
        for j in range(3):
            assert abs(Kdev[i, j] - k(X[i], i, X[j], j)) < 1e-13
    assert abs(model.objective_function() - nlml) < 1e-10 * max(1.0, abs(nlml))
    T4 = np.array([xs[t] + [float(f)] for t in range(2) for f in range(3)])
    mu, var = w.predict(T4)
    for r, (t, f) in enumerate((t, f) for t in range(2) for f in range(3)):
        ks = [k(xs[t], f, X[j], j) for j in range(3)]
        m_want = sum(ks[j] * alpha[j] for j in range(3))
        v_want = k(xs[t], f, xs[t], f) - sum(ks[i] * Kinv[i][j] * ks[j] for i in range(3) for j in range(3)) + noise[f]
        assert abs(mu[r, 0] - m_want) < 1e-12 * max(1.0, abs(m_want)), (r, mu[r, 0], m_want)
        assert abs(var[r, 0] - v_want) < 1e-12 * max(1.0, v_want), (r, var[r, 0], v_want)
    C = w.predict_covariance(T4)
    for a, (ta, fa) in enumerate((t, f) for t in range(2) for f in range(3)):
        for b, (tb, fb) in enumerate((t, f) for t in range(2) for f in range(3)):
            ka = [k(xs[ta], fa, X[j], j) for j in range(3)]
            kbv = [k(xs[tb], fb, X[j], j) for j in range(3)]
            want = k(xs[ta], fa, xs[tb], fb) - sum(ka[i] * Kinv[i][j] * kbv[j] for i in range(3) for j in range(3))
            if a == b:
                want += noise[fa]
            want = max(want, 1e-10)                                   # emukit: np.clip(cov, 1e-10, inf)
            assert abs(C[a, b] - want) < 1e-12 * max(1.0, abs(want)), (a, b, C[a, b], want)


def test_gptrainers_flow_on_twelve_bundled_datasets(gpcore_mod):
    """BASELINE configs[0] widened: the reference's GPTrainers.py flow on twelve bundled data sets of field 0
    (tests/golden/gp_datasets.npz) against the numbers the reference published for each
    (Data/TrajectoriesAndEstimates/GPResults/MSE_*.txt, produced with real GPy / emukit): the single- and multi-fidelity
    RMSEs inside a 1e-4 band (measured: <= 5e-6), the single-fidelity covariance-weighted MSEs -- which involve the full
    2000 x 2000 posterior covariance -- inside 1e-3 (measured: 4 digits), on at least eleven of the twelve.  They depend
    on where the reference's unseeded optimisers stopped, so this is a band, not a pin; the NIGP number moves with the
    reference's unseeded restarts and is only required to stay inside 5 %.  This is the evidence at the GPy / emukit
    boundary that does not pass through oracle/gp_oracle.py."""
    import importlib.util
    import os
    from conftest import ROOT
    spec = importlib.util.spec_from_file_location("gptrainers_flow", os.path.join(ROOT, "examples", "gptrainers_flow.py"))
    flow = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(flow)
    g = golden("gp_datasets.npz")
    cn = [str(c) for c in g["columns"]]
    Lsw = g["field0_Lsw"]
    field = dict(L=Lsw[0], s=Lsw[1], w=Lsw[2:], p=g["field0_p"])
    inside, report = 0, []
    for name in [str(n) for n in g["names"]]:
        d = g["data_" + name]
        d = d[d[:, 0] < 3600]                                        # GPTrainers.py:37
        cols = {c: d[:, i] for i, c in enumerate(cn)}
        np.random.seed(0)
        rm, wm = flow.run(cols, field=field, verbose=False)
        pub = dict(zip(("mf", "sf", "nisf", "sfTP", "w_mf", "w_sf", "w_nisf", "w_sfTP"), g["pub_" + name]))
        dev = {kk: abs(rm[kk] - pub[kk]) / pub[kk] for kk in ("mf", "sf", "sfTP", "nisf")}
        dev.update({"w_" + kk: abs(float(wm[kk]) - pub["w_" + kk]) / pub["w_" + kk] for kk in ("sf", "sfTP")})
        assert all(np.isfinite(rm[kk]) for kk in rm), (name, rm)
        ok = dev["mf"] < 1e-4 and dev["sf"] < 1e-4 and dev["sfTP"] < 1e-4 and dev["w_sf"] < 1e-3 and dev["w_sfTP"] < 1e-3 \
            and dev["nisf"] < 5e-2
        inside += bool(ok)
        report.append((name, ok, {k2: float("%.2e" % v2) for k2, v2 in dev.items()}))
    for r in report:
        print(r)
    # one of the twelve (T0_0.1) is a data set on which the reference's own multi-fidelity fit diverged (published RMSE
    # 34.9 against ~7 everywhere else): the optimiser's path, not the arithmetic, decides that number
    assert inside >= 11, report
