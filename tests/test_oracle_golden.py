"""CPU: the NumPy restatement (oracle/gp_oracle.py) against the golden vectors produced by the
reference's own NIGP.py (tests/golden/nigp_*.npz, generator oracle/make_golden.py), against its
own frozen outputs, and refit-loop vs Schur-complement information gain."""
import numpy as np
import pytest

from conftest import golden, normwise
from oracle import gp_oracle as go


@pytest.mark.parametrize("gram", [True, False])
def test_nigp_demo_predict_matches_reference(gram):
    g = golden("nigp_demo.npz")
    mean, var = go.nigp_predict(g["X"], g["y"], g["lengthscales"], float(g["sigma_f"]), float(g["sigma_y"]),
                                g["noise_diag"], g["Xs"], gram=gram)
    tol = 1e-12 if gram else 1e-9
    assert normwise(mean, g["mean"]) < tol
    assert normwise(var, g["var"], float(g["sigma_f"])) < tol
    _, cov = go.nigp_predict(g["X"], g["y"], g["lengthscales"], float(g["sigma_f"]), float(g["sigma_y"]),
                             g["noise_diag"], g["Xs"], return_cov=True, gram=gram)
    assert normwise(cov, g["cov"]) < tol
    _, var_in = go.nigp_predict(g["X"], g["y"], g["lengthscales"], float(g["sigma_f"]), float(g["sigma_y"]),
                                g["noise_diag"], g["Xs"], Xs_input_noise=g["sigma_x"], gram=gram)
    assert normwise(var_in, g["var_in"], float(g["sigma_f"])) < tol


def test_nigp_demo_kernel_gradients_nlml():
    g = golden("nigp_demo.npz")
    K = go.SE_ARD_kernel(g["X"], g["Xs"], g["lengthscales"], float(g["sigma_f"]))
    assert normwise(K, g["K"]) < 1e-13
    fm, grads = go.compute_post_mean_and_gradients(g["X"], g["y"], g["lengthscales"], float(g["sigma_f"]),
                                                   float(g["sigma_y"]), g["noise_diag"])
    assert normwise(fm, g["f_mean_train"]) < 1e-11
    assert normwise(grads, g["grads"]) < 1e-11
    nlml = go.nigp_nlml(g["log_hyp"], g["X"], g["y"], g["grads"])
    assert abs(nlml - float(g["nlml"])) < 1e-9 * abs(float(g["nlml"]))


def test_nigp_field_matches_reference():
    g, d = golden("nigp_field.npz"), golden("field_data.npz")
    fm, grads = go.compute_post_mean_and_gradients(d["Xh"], d["y"], g["ls"], float(g["sigma_f"]), float(g["sigma_y"]))
    assert normwise(fm, g["f_mean_train"]) < 1e-10
    assert normwise(grads, g["grads"]) < 1e-10
    assert abs(go.nigp_nlml(g["log_hyp"], d["Xh"], d["y"], g["grads"]) - float(g["nlml"])) < 1e-8
    assert abs(go.nigp_nlml(g["log_hyp"], d["Xh"], d["y"], g["grads"], 0.01 * np.ones(len(d["y"])))
               - float(g["nlml_extra"])) < 1e-8
    mean, var = go.nigp_predict(d["Xh"], d["y"], g["ls"], float(g["sigma_f"]), float(g["sigma_y"]), g["noise_diag"],
                                d["test"])
    assert normwise(mean, g["mean"]) < 1e-10
    assert normwise(var, g["var"], float(g["sigma_f"])) < 1e-10
    _, var_in = go.nigp_predict(d["Xh"], d["y"], g["ls"], float(g["sigma_f"]), float(g["sigma_y"]), g["noise_diag"],
                                d["test"], Xs_input_noise=g["sigma_x"])
    assert normwise(var_in, g["var_in"], float(g["sigma_f"])) < 1e-10


def test_restatement_frozen_outputs():
    g, d = golden("gp_oracle.npz"), golden("field_data.npz")
    gp = go.SFGP(d["Xh"], d["y"], g["sf_params"])
    mu, var = gp.predict(d["test"])
    assert normwise(mu, g["mu_sf"]) < 1e-11 and normwise(var, g["var_sf"]) < 1e-11
    assert abs(gp.f.nlml - float(g["nlml_sf"])) < 1e-8
    mf = go.MFGP(g["X4"], g["y4"], g["mf_params"], F=3)
    t4 = np.hstack([d["test"], 2 * np.ones((len(d["test"]), 1))])
    mu, var = mf.predict(t4)
    assert normwise(mu, g["mu_mf"]) < 1e-11 and normwise(var, g["var_mf"]) < 1e-11


def test_gram_and_direct_distances_agree():
    d = golden("field_data.npz")
    a = go.k_stationary(d["Xh"][:200], d["test"][:300], 4.0, np.array([2.0, 3.0, 2.5]), gram=True)
    b = go.k_stationary(d["Xh"][:200], d["test"][:300], 4.0, np.array([2.0, 3.0, 2.5]), gram=False)
    assert np.max(np.abs(a - b)) < 1e-12


def test_ar1_kernel_structure():
    rng = np.random.default_rng(1)
    X4 = np.hstack([rng.uniform(0, 5, (60, 3)), rng.integers(0, 3, (60, 1)).astype(float)])
    v = [2.0, 1.0, 0.5]; ls = np.array([[1, 2, 1.5], [0.5, 1, 1], [2, 2, 2]], float); rho = [0.8, 1.2]
    K = go.k_ar1(X4, X4, v, ls, rho, same=True)
    assert np.allclose(K, K.T)
    assert np.min(np.linalg.eigvalsh(K + 1e-9 * np.eye(60))) > 0
    assert np.allclose(np.diag(K), go.k_ar1_diag(X4, v, rho))
    # fidelity-0 block is the base kernel, fidelity-1 block is rho0^2 k0 + k1
    i0 = X4[:, 3] == 0
    assert np.allclose(K[np.ix_(i0, i0)], go.k_stationary(X4[i0, :3], X4[i0, :3], v[0], ls[0], same=True))
    i1 = X4[:, 3] == 1
    want = rho[0] ** 2 * go.k_stationary(X4[i1, :3], X4[i1, :3], v[0], ls[0], same=True) + \
        go.k_stationary(X4[i1, :3], X4[i1, :3], v[1], ls[1], same=True)
    assert np.allclose(K[np.ix_(i1, i1)], want)


def test_ig_schur_equals_refit_loop():
    rng = np.random.default_rng(3)
    X = rng.uniform(0, 6, (80, 3)); y = np.sin(X[:, 0])
    p = np.array([2.0, 1.5, 2.0, 1.0, 0.05])
    gp = go.SFGP(X, y, p)
    for k in (1, 5, 32):
        Xc = rng.uniform(0, 6, (k, 3))
        a = go.ig_seq_sf_refit(gp, Xc, first_preadded=True)
        b = go.ig_seq_schur(gp, Xc, p[-1], p[-1], p[-1], first_preadded=True)
        assert abs(a - b) < 1e-10 * max(1, abs(a))
        grid = rng.uniform(0, 6, (30, 3))
        c = go.ig_logdet_refit(gp, grid, Xc)
        d = go.ig_logdet_schur(gp, grid, Xc, p[-1], p[-1])
        assert abs(c - d) < 1e-9 * max(1, abs(c))
    X4 = np.hstack([X, rng.integers(0, 3, (80, 1)).astype(float)])
    mp = np.array([3.0, 2.5, 3.5, 3.0, 1.0, 1.5, 2.0, 2.0, 0.5, 1.0, 1.5, 1.5, 0.9, 1.1, 0.08, 0.04, 0.02])
    mf = go.MFGP(X4, y, mp, F=3)
    Xc4 = np.hstack([rng.uniform(0, 6, (9, 3)), rng.integers(0, 3, (9, 1)).astype(float)])
    a = go.ig_seq_mf_refit(mf, Xc4, mp[-1], pred_fid=0)
    Xp = Xc4.copy(); Xp[:, 3] = 0
    b = go.ig_seq_schur(mf, Xc4, mf.noise_of(Xc4), mf.noise_of(Xp), mp[-1], Xpred=Xp)
    assert abs(a - b) < 1e-10 * max(1, abs(a))


def test_published_rmse_sanity_band():
    """GPTrainers.py config 1 sanity: an SF-GP with plausible fixed hypers on the bundled dataset
    must land inside the loose band of the published per-file RMSE (results.csv, ~5.25 for
    fieldMeas_0_T0_0) -- not a 1e-9 pin, just a guard against a wrong data layout."""
    d = golden("field_data.npz")
    gp = go.SFGP(d["Xh"], d["y"], np.array([4.0, 2.0, 3.0, 2.5, 0.05]))
    mu, _ = gp.predict(d["Xh"][:50])
    assert np.sqrt(np.mean((mu[:, 0] - d["y"][:50]) ** 2)) < 1.0


def _golden_paths(g):
    paths = []
    eo, po = g["edge_off"], g["prim_off"]
    for c in range(len(eo) - 1):
        path = []
        for e in range(eo[c], eo[c + 1]):
            xy = g["edge_xy"][e]
            path.append((xy[:2], xy[2:], [tuple(p) for p in g["prims"][po[e]:po[e + 1]]]))
        paths.append(path)
    return paths


@pytest.mark.parametrize("dense", [0, 1])
@pytest.mark.parametrize("wv", [0, 1])
def test_trajectory_restatement_matches_reference_planner(dense, wv):
    """oracle/traj_oracle.py against outputs of the reference's own GraceAgent.pathToTrajPoints
    (tests/golden/traj_paths.npz, generated by importing /root/reference/GraceRIGV3.py)."""
    from oracle import traj_oracle as to
    g = golden("traj_paths.npz")
    pts, off = g["pts_d%d_v%d" % (dense, wv)], g["off_d%d_v%d" % (dense, wv)]
    for c, path in enumerate(_golden_paths(g)):
        got = to.path_to_traj_points(path, float(g["variance_rate"]), float(g["meas_rate"]), dense=bool(dense), with_var=bool(wv))
        want = pts[off[c]:off[c + 1]]
        assert got.shape == want.shape and np.max(np.abs(got - want), initial=0.0) < 1e-12, c


def test_getEID_host_logic_matches_reference():
    """gpcore.eid.getEID (host part of the EID map) over oracle-backed adapters against golden values produced by
    the reference's own exploreSimSettings.getEID (tests/golden/eid.npz)."""
    import types
    import __graft_entry__ as entry
    entry.setup_path()
    from gpcore.eid import getEID
    from oracle import gp_oracle as go
    g = golden("eid.npz")
    sf = go.SFGP(g["Xh"], g["y"][:, None], g["sf_params"], gram=False)
    mf = go.MFGP(g["X4"], g["y"][:, None], g["mf_params"], F=3, gram=False)
    sfa = types.SimpleNamespace(predict=lambda X: sf.predict(np.asarray(X, float)),
                                kern=types.SimpleNamespace(variance=np.array([g["sf_params"][0]])),
                                Gaussian_noise=types.SimpleNamespace(variance=np.array([g["sf_params"][-1]])))
    mfa = types.SimpleNamespace(predict=lambda X: mf.predict(np.asarray(X, float)),
                                gpy_model=types.SimpleNamespace(param_array=g["mf_params"]))
    for auto in (0, 1):
        E, grid = getEID(sfa, g["WS"], float(g["mD"]), auto=auto)
        assert np.array_equal(grid, g["grid"])
        assert np.max(np.abs(E - g["sf_auto%d" % auto])) < 1e-12 * np.max(g["sf_auto%d" % auto])
        E, _ = getEID(mfa, g["WS"], float(g["mD"]), emu=True, auto=auto)
        assert np.max(np.abs(E - g["mf_auto%d" % auto])) < 1e-12 * np.max(g["mf_auto%d" % auto])
    E, _ = getEID(sfa, g["WS"], float(g["mD"]), alpha=0.3)
    assert np.max(np.abs(E - g["sf_alpha03"])) < 1e-12 * np.max(g["sf_alpha03"])
    # experiment variant: negative variances are replaced by the prior variance, sqrt without abs
    neg = types.SimpleNamespace(predict=lambda X: (np.zeros((len(X), 1)), np.where(np.arange(len(X))[:, None] % 7 == 0, -1e-6, 0.5)),
                                kern=sfa.kern, Gaussian_noise=sfa.Gaussian_noise)
    E, ss = getEID(neg, g["WS"], float(g["mD"]), variant="exp", alpha=0.2, default_grid=g["grid"][:100])
    assert np.all(np.isfinite(E)) and abs(E.sum() - 1) < 1e-12 and ss.shape == (100, 3)
    E, _ = getEID(neg, g["WS"], float(g["mD"]))                    # simulation variant: uniform map
    assert np.allclose(E, 1.0 / E.shape[0])


def test_scale_fixtures_were_generated_on_the_seeded_inputs():
    """tests/golden/scale_oracle.npz and nigp_8192.npz freeze oracle / reference outputs at the BASELINE sizes; the GPU
    tests use them only when the checksum of the regenerated inputs matches.  Check that it does (else those tests fall
    back to minutes of live oracle work on the GPU box)."""
    import hashlib
    import scale_cases as sc
    import bench
    g = golden("scale_oracle.npz")
    X4, y, p, Xs4 = sc.configs1_inputs()
    assert str(g["c1_sha"]) == sc.sha(X4, y, p, Xs4)
    X4, y, p, cands, grid4 = sc.configs3_inputs()
    assert str(g["c3_sha"]) == sc.sha(X4, y, p, grid4, *cands)
    X, y, p, Xs = sc.sf16384_inputs()
    assert str(g["s16_sha"]) == sc.sha(X, y, p, Xs)
    n = golden("nigp_8192.npz")
    X4, y = bench.make_train(int(n["N"]), 3)
    X = np.ascontiguousarray(X4[:, :3])
    assert hashlib.sha256(X.tobytes() + y.tobytes()).hexdigest() == str(n["train_sha256"])
    # the literal refit loop and the Schur form agree at N = 4096 as well (CPU vs CPU, frozen)
    k = len(g["c3_seq_loop"])
    assert normwise(g["c3_seq_schur"][:k], g["c3_seq_loop"], 1.0) < 1e-10


def test_restated_nigp_fit_lands_on_the_reference_fit():
    """oracle ``nigp_fit`` (NIGP.py:191-260 restated) against the reference's own fit (nigp_fit.npz), both distance
    formulations, within 10 x the recorded CPU-vs-CPU spread."""
    g = golden("nigp_fit.npz")
    tol = max(1e-6, 10.0 * float(g["spread_params"]))
    for gram in (True, False):
        np.random.seed(0)
        np.random.randn(40, 1)
        np.random.randn(40)
        p, _, nd, _ = go.nigp_fit(g["X"], g["y"], int(g["n_restarts"]), int(g["iters"]), int(g["maxiter_opt"]), gram=gram)
        assert np.max(np.abs(p - g["params"]) / np.abs(g["params"])) < tol
        assert normwise(nd, g["noise_diag"]) < 100 * tol
