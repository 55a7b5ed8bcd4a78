"""GPU parity: the CUDA path (through the C ABI / the reference-facing Python surface) against
the oracle and the committed golden vectors.

Tolerance (BASELINE.json north_star): 1e-9 relative in FP64, measured NORMWISE --
max|gpu - ref| <= 1e-9 * max(|ref|, kernel variance) -- because var = k** - |L^-1 k*|^2 cancels
(SURVEY.md section 7).  Reference values come from (a) the reference's own NIGP.py
(tests/golden/nigp_*.npz, pinned) and (b) the NumPy restatement of the GPy/emukit arithmetic
(unpinned third-party boundary), evaluated with direct distances (`gram=False`, the same
formulation the kernels use) and with GPy's Gram-trick distances (`gram=True`, the frozen golden arrays), both at
1e-9: every assert of this file and of tests/test_gpu_int8.py that round 1 had loosened to 1e-8 .. 1e-7 now holds the
north star's 1e-9 (measured values are printed by `measured()` and logged; the worst is 7e-10, an FP64 mean at
cond(K) ~ 1e7).
"""
import numpy as np
import pytest

from conftest import golden, measured, normwise

pytestmark = pytest.mark.gpu

TOL = 1e-9
GRAM_TOL = TOL     # round 1 allowed 2e-8 against the Gram-trick goldens; the measured spread is <= 4e-11


@pytest.fixture(scope="module")
def gpcore_mod(built_lib):
    import gpcore
    return gpcore


@pytest.fixture(scope="module")
def go():
    from oracle import gp_oracle
    return gp_oracle


MF_PARAMS = np.array([3.0, 2.5, 3.5, 3.0, 1.0, 1.5, 2.0, 2.0, 0.5, 1.0, 1.5, 1.5, 0.9, 1.1, 0.08, 0.04, 0.02])
SF_PARAMS = np.array([4.0, 2.0, 3.0, 2.5, 0.05])


def synth(rng, N, F):
    X = rng.uniform([0, 0, 0], [10, 20, 10], (N, 3))
    f = rng.integers(0, F, (N, 1)).astype(float) if F > 1 else np.zeros((N, 1))
    y = np.sin(X[:, 0]) * np.cos(0.3 * X[:, 1]) + 0.2 * f[:, 0] + 0.05 * rng.standard_normal(N)
    return np.hstack([X, f]), y


# ---------------------------------------------------------------------------------------------
# factorisation pieces
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("N", [1, 5, 127, 128, 129, 300, 709, 1500])
def test_factor_chol_alpha_logdet(gpcore_mod, go, N):
    L_ = gpcore_mod._lib
    rng = np.random.default_rng(N)
    X4, y = synth(rng, N, 1)
    core = gpcore_mod.GPCore(L_.KIND_SF_RBF, 1, 0)
    core.set_hypers(SF_PARAMS, 1e-8)
    core.set_data(X4, y)
    nlml, logdet = core.factor()
    ref = go.SFGP(X4[:, :3], y, SF_PARAMS, gram=False)
    assert normwise(core.chol(), ref.f.L) < 1e-11
    assert normwise(core.alpha(), ref.f.alpha) < TOL
    assert abs(logdet - ref.f.logdet) < TOL * max(1.0, abs(ref.f.logdet))
    assert abs(nlml - ref.f.nlml) < TOL * max(1.0, abs(ref.f.nlml))
    Linv = core.linv()
    assert normwise(Linv @ ref.f.L, np.eye(N), 1.0) < 1e-10
    assert core.padded_n() % 128 == 0 and core.padded_n() >= N
    core.close()


def test_not_positive_definite_is_recoverable(gpcore_mod):
    L_ = gpcore_mod._lib
    X4 = np.zeros((10, 4))            # ten identical points, zero noise, zero jitter -> singular
    core = gpcore_mod.GPCore(L_.KIND_SF_RBF, 1, 0)
    core.set_hypers(np.array([1.0, 1, 1, 1, 0.0]), 0.0)
    core.set_data(X4, np.ones(10))
    with pytest.raises(np.linalg.LinAlgError):
        core.factor()
    core.set_hypers(np.array([1.0, 1, 1, 1, 0.1]), 1e-8)   # and the handle is still usable
    nlml, _ = core.factor()
    assert np.isfinite(nlml)
    with pytest.raises(ValueError):
        core.set_hypers(np.ones(4), 0.0)
    core.close()


def test_state_errors(gpcore_mod):
    L_ = gpcore_mod._lib
    core = gpcore_mod.GPCore(L_.KIND_SF_RBF, 1, 0)
    with pytest.raises(gpcore_mod.GpcoreError):
        core.predict(np.zeros((3, 4)), 0)
    core.close()


# ---------------------------------------------------------------------------------------------
# kernel matrices
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", ["rbf", "mat32"])
def test_kernel_matrix_sf_and_mf(gpcore_mod, go, kind):
    L_ = gpcore_mod._lib
    rng = np.random.default_rng(2)
    A, _ = synth(rng, 70, 3)
    B, _ = synth(rng, 45, 3)
    ok = go.KIND_RBF if kind == "rbf" else go.KIND_MAT32
    core = gpcore_mod.GPCore(L_.KIND_SF_RBF if kind == "rbf" else L_.KIND_SF_MAT32, 1, 0)
    core.set_hypers(SF_PARAMS, 0.0)
    K = core.kernel_matrix(A, B)
    assert normwise(K, go.k_stationary(A[:, :3], B[:, :3], 4.0, SF_PARAMS[1:4], ok, gram=False)) < 1e-13
    assert normwise(K, go.k_stationary(A[:, :3], B[:, :3], 4.0, SF_PARAMS[1:4], ok, gram=True)) < 1e-12
    core.close()
    core = gpcore_mod.GPCore(L_.KIND_MF_AR1_RBF if kind == "rbf" else L_.KIND_MF_AR1_MAT32, 3, 0)
    core.set_hypers(MF_PARAMS, 0.0)
    v, ls, rho, _ = go.split_mf_params(MF_PARAMS, 3)
    assert normwise(core.kernel_matrix(A, B), go.k_ar1(A, B, v, ls, rho, ok, gram=False)) < 1e-13
    assert normwise(core.kernel_matrix(A), go.k_ar1(A, A, v, ls, rho, ok, gram=False, same=True)) < 1e-13
    core.close()


# ---------------------------------------------------------------------------------------------
# NIGP -- pinned by the reference's own module (golden fixtures)
# ---------------------------------------------------------------------------------------------
def test_nigp_module_functions_demo(gpcore_mod):
    from gpcore import nigp
    g = golden("nigp_demo.npz")
    K = nigp.SE_ARD_kernel(g["X"], g["Xs"], g["lengthscales"], float(g["sigma_f"]))
    assert normwise(K, g["K"]) < 1e-12
    fm, grads = nigp.compute_post_mean_and_gradients(g["X"], g["y"], g["lengthscales"], float(g["sigma_f"]),
                                                     float(g["sigma_y"]), g["noise_diag"])
    assert fm.shape == (40,) and grads.shape == (40, 1)
    assert normwise(fm, g["f_mean_train"]) < TOL and normwise(grads, g["grads"]) < TOL
    nlml = nigp.neg_log_marginal_likelihood(g["log_hyp"], g["X"], g["y"], g["grads"])
    assert abs(nlml - float(g["nlml"])) < TOL * abs(float(g["nlml"]))
    assert nigp.safe_obj(g["log_hyp"], g["X"], g["y"], g["grads"], None) == nlml


def test_nigp_class_predict_demo(gpcore_mod):
    from gpcore.nigp import NIGP
    g = golden("nigp_demo.npz")
    m = NIGP(verbose=False)
    m.lengthscales_, m.sigma_f_, m.sigma_y_, m.sigma_x_ = g["lengthscales"], float(g["sigma_f"]), float(g["sigma_y"]), g["sigma_x"]
    m.X_train_, m.y_train_, m.noise_diag_train_ = g["X"], g["y"], g["noise_diag"]
    assert np.allclose(m.get_params(), g["params"])
    mean, var = m.predict(g["Xs"])
    sf = float(g["sigma_f"])
    assert normwise(mean, g["mean"]) < TOL and normwise(var, g["var"], sf) < TOL
    mean2, cov = m.predict(g["Xs"], return_cov=True)
    assert normwise(cov, g["cov"], sf) < TOL and np.allclose(mean2, mean)
    assert normwise(m.predict(g["Xs"], return_var=False), g["mean"]) < TOL
    _, var_in = m.predict(g["Xs"], Xs_input_noise=g["sigma_x"])
    assert normwise(var_in, g["var_in"], sf) < TOL
    _, cov_in = m.predict(g["Xs"][:50], Xs_input_noise=np.full((50, 1), 0.1), return_cov=True)
    assert normwise(cov_in, g["cov_in"], sf) < TOL
    with pytest.raises(ValueError):
        m.predict(g["Xs"], Xs_input_noise=0.1)        # a scalar raises in the reference too (NIGP.py:313-319)


def test_nigp_field_dataset(gpcore_mod):
    from gpcore import nigp
    g, d = golden("nigp_field.npz"), golden("field_data.npz")
    sf = float(g["sigma_f"])
    fm, grads = nigp.compute_post_mean_and_gradients(d["Xh"], d["y"], g["ls"], sf, float(g["sigma_y"]))
    assert normwise(fm, g["f_mean_train"]) < TOL and normwise(grads, g["grads"]) < TOL
    assert abs(nigp.neg_log_marginal_likelihood(g["log_hyp"], d["Xh"], d["y"], g["grads"]) - float(g["nlml"])) \
        < TOL * max(1.0, abs(float(g["nlml"]))) * 10
    assert abs(nigp.neg_log_marginal_likelihood(g["log_hyp"], d["Xh"], d["y"], g["grads"], 0.01 * np.ones(len(d["y"])))
               - float(g["nlml_extra"])) < TOL * 100
    m = nigp.NIGP(verbose=False)
    m.lengthscales_, m.sigma_f_, m.sigma_y_, m.sigma_x_ = g["ls"], sf, float(g["sigma_y"]), g["sigma_x"]
    m.X_train_, m.y_train_, m.noise_diag_train_ = d["Xh"], d["y"], g["noise_diag"]
    mean, var = m.predict(d["test"])
    assert normwise(mean, g["mean"]) < TOL and normwise(var, g["var"], sf) < TOL
    _, cov = m.predict(d["test_sub"], return_cov=True)
    assert normwise(cov, g["cov_sub"], sf) < TOL
    _, var_in = m.predict(d["test"], Xs_input_noise=g["sigma_x"])
    assert normwise(var_in, g["var_in"], sf) < TOL
    _, var_in2 = m.predict(d["test"], Xs_input_noise=np.tile(g["sigma_x"], (len(d["test"]), 1)))
    assert np.allclose(var_in2, var_in, rtol=1e-13, atol=0)


def test_nigp_fit_runs_and_improves(gpcore_mod):
    """fit() drives SciPy's L-BFGS-B with the device NLML: the optimised NLML must not be worse
    than the starting point, attributes have the reference's shapes."""
    from gpcore import nigp
    g = golden("nigp_demo.npz")
    np.random.seed(0)
    m = nigp.NIGP(n_restarts=1, iters=2, verbose=False).fit(g["X"], g["y"], maxiter_opt=30)
    assert m.lengthscales_.shape == (1,) and m.sigma_x_.shape == (1,) and m.noise_diag_train_.shape == (40,)
    start = nigp.NIGP._initial_log_hypers(g["X"], g["y"])
    lh = np.log(np.concatenate([m.lengthscales_, [m.sigma_f_, m.sigma_y_], m.sigma_x_]))
    zeros = np.zeros_like(g["X"])
    assert nigp.neg_log_marginal_likelihood(lh, g["X"], g["y"], zeros) <= \
        nigp.neg_log_marginal_likelihood(start, g["X"], g["y"], zeros) + 1e-6
    mean, var = m.predict(g["Xs"])
    assert np.all(np.isfinite(mean)) and np.all(var >= 1e-12)


def test_nigp_fit_reproduces_the_reference_fit(gpcore_mod):
    """``NIGP.fit`` parity (``NIGP.py:191-260``): with the reference's own scheme (numerical gradients,
    ``analytic_grad=False``), the ``__main__`` protocol (n_restarts = 2, iters = 10) and NumPy's global generator in the
    state the reference's run had, the device-backed fit lands on the hypers the reference's NIGP.py fitted
    (tests/golden/nigp_fit.npz).  Tolerance: 1e-6, or 10 x the CPU-vs-CPU spread the golden records (L-BFGS-B
    differentiates with a 1e-8 step, so rounding noise in the objective moves the fitted point; two CPU formulations of
    the same objective already differ by ``spread_params``).  The analytic-gradient path (the default here) must reach
    an objective at least as low."""
    from gpcore import nigp
    g = golden("nigp_fit.npz")
    X, y = g["X"], g["y"]
    tol = max(1e-6, 10.0 * float(g["spread_params"]))

    def seeded():
        np.random.seed(0)
        np.random.randn(40, 1)
        np.random.randn(40)

    seeded()
    m = nigp.NIGP(n_restarts=int(g["n_restarts"]), iters=int(g["iters"]), verbose=False, analytic_grad=False)
    m.fit(X, y, maxiter_opt=int(g["maxiter_opt"]))
    rel = np.abs(m.get_params() - g["params"]) / np.abs(g["params"])
    print("NIGP.fit vs reference: max rel %.3e (tolerance %.1e, CPU-vs-CPU spread %.1e)" % (rel.max(), tol, float(g["spread_params"])))
    assert rel.max() < tol, (m.get_params(), g["params"])
    assert normwise(m.noise_diag_train_, g["noise_diag"]) < 100 * tol
    lh = np.log(np.concatenate([m.lengthscales_, [m.sigma_f_, m.sigma_y_], m.sigma_x_]))
    zeros = np.zeros_like(X)
    f_fd = nigp.neg_log_marginal_likelihood(lh, X, y, zeros, m.noise_diag_train_)
    assert abs(f_fd - float(g["nlml_at_fit"])) < 1e-6 * max(1.0, abs(float(g["nlml_at_fit"])))
    # analytic gradients: same alternation, better-conditioned search -- never a worse objective at its own fitted point
    seeded()
    a = nigp.NIGP(n_restarts=int(g["n_restarts"]), iters=int(g["iters"]), verbose=False, analytic_grad=True)
    a.fit(X, y, maxiter_opt=int(g["maxiter_opt"]))
    la = np.log(np.concatenate([a.lengthscales_, [a.sigma_f_, a.sigma_y_], a.sigma_x_]))
    f_an = nigp.neg_log_marginal_likelihood(la, X, y, zeros, a.noise_diag_train_)
    assert f_an <= float(g["nlml_at_fit"]) + 1e-4, (f_an, float(g["nlml_at_fit"]))


# ---------------------------------------------------------------------------------------------
# SF / MF models through the GPy / emukit mirror
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", ["rbf", "mat32"])
def test_sfgp_predict_field(gpcore_mod, go, kind):
    from gpcore.GPy.kern import RBF, Matern32
    from gpcore.GPy.models import GPRegression
    d, g = golden("field_data.npz"), golden("gp_oracle.npz")
    kcls = RBF if kind == "rbf" else Matern32
    gp = GPRegression(d["Xh"], d["y"][:, None], kcls(input_dim=3, ARD=True))
    gp.param_array[:] = SF_PARAMS
    mu, var = gp.predict(d["test"])
    assert mu.shape == (2000, 1) and var.shape == (2000, 1)
    ok = go.KIND_RBF if kind == "rbf" else go.KIND_MAT32
    ref = go.SFGP(d["Xh"], d["y"], SF_PARAMS, kind=ok, gram=False)
    mu0, var0 = ref.predict(d["test"])
    assert normwise(mu, mu0) < TOL and normwise(var, var0, 4.0) < TOL
    gmu, gvar = (g["mu_sf"], g["var_sf"]) if kind == "rbf" else (g["mu_32"], g["var_32"])
    assert measured("sf_vs_gram_golden_mean_" + kind, normwise(mu, gmu), GRAM_TOL) < GRAM_TOL   # Gram-trick CPU formulation
    assert measured("sf_vs_gram_golden_var_" + kind, normwise(var, gvar, 4.0), GRAM_TOL) < GRAM_TOL
    assert abs(gp.objective_function() - ref.f.nlml) < TOL * abs(ref.f.nlml)
    mu2, cov = gp.predict(d["test_sub"], full_cov=1)
    _, cov0 = ref.predict(d["test_sub"], full_cov=True)
    assert cov.shape == (250, 250) and normwise(cov, cov0, 4.0) < TOL
    assert np.allclose(cov, cov.T, rtol=0, atol=0)
    # the diagonal of the full covariance is the marginal variance (before the 1e-15 clip)
    _, var_sub = gp.predict(d["test_sub"])
    # (the diagonal path runs on the INT8 tensor cores by default, the full covariance in FP64 DMMA:
    #  they agree to the ~1e-11 the 7-digit splitting carries, two orders inside the parity tolerance)
    assert normwise(np.diag(cov), var_sub[:, 0], 4.0) < 1e-10
    # noise-free latent prediction
    _, lat = gp.predict(d["test_sub"], include_likelihood=False)
    assert normwise(lat + SF_PARAMS[-1], var_sub, 4.0) < 1e-12


def test_sfgp_set_xy_copy_and_noise_assignment(gpcore_mod, go):
    from gpcore.GPy.kern import RBF
    from gpcore.GPy.models import GPRegression
    rng = np.random.default_rng(5)
    X4, y = synth(rng, 200, 1)
    gp = GPRegression(X4[:100, :3], y[:100, None], RBF(3, variance=4.0, lengthscale=[2, 3, 2.5], ARD=True))
    gp.Gaussian_noise.variance = 0.05
    c = gp.copy()
    gp.set_XY(X4[:, :3], y[:, None])
    Xs = rng.uniform(0, 10, (64, 3))
    mu, var = gp.predict(Xs)
    ref = go.SFGP(X4[:, :3], y, SF_PARAMS, gram=False)
    mu0, var0 = ref.predict(Xs)
    assert normwise(mu, mu0) < TOL and normwise(var, var0, 4.0) < TOL
    muc, _ = c.predict(Xs)                       # the copy still holds the first 100 points
    refc = go.SFGP(X4[:100, :3], y[:100], SF_PARAMS, gram=False)
    assert normwise(muc, refc.predict(Xs)[0]) < TOL
    # 1-D, non-ARD model as in informationGainTest.py:22-33
    x1 = np.linspace(0, 5, 30)[:, None]
    m1 = GPRegression(x1, np.sin(x1), RBF(input_dim=1, variance=0.74, lengthscale=1.57))
    m1.Gaussian_noise.variance = 0.00104
    r1 = go.SFGP(np.hstack([x1, np.zeros((30, 2))]), np.sin(x1), np.array([0.74, 1.57, 1.0, 1.0, 0.00104]), gram=False)
    q = np.linspace(-1, 6, 40)[:, None]
    a, b = m1.predict(q)
    a0, b0 = r1.predict(np.hstack([q, np.zeros((40, 2))]))
    assert normwise(a, a0) < TOL and normwise(b, b0, 0.74) < TOL


@pytest.mark.parametrize("noise_mode", ["mixed", "single"])
def test_mfgp_predict_field(gpcore_mod, go, noise_mode):
    from gpcore.GPy.kern import RBF
    from gpcore.GPy.likelihoods import Gaussian
    from gpcore.emukit.multi_fidelity.kernels import LinearMultiFidelityKernel
    from gpcore.emukit.multi_fidelity.models import GPyLinearMultiFidelityModel
    from gpcore.emukit.model_wrappers.gpy_model_wrappers import GPyMultiOutputWrapper
    d, g = golden("field_data.npz"), golden("gp_oracle.npz")
    params = MF_PARAMS if noise_mode == "mixed" else MF_PARAMS[:15]
    k = LinearMultiFidelityKernel([RBF(3, ARD=True), RBF(3, ARD=True), RBF(3, ARD=True)])
    lik = None if noise_mode == "mixed" else Gaussian()
    m = GPyLinearMultiFidelityModel(g["X4"], g["y4"][:, None], k, likelihood=lik, n_fidelities=3)
    w = GPyMultiOutputWrapper(m, 3, n_optimization_restarts=1)
    w.set_data(g["X4"], g["y4"][:, None])
    m.param_array[:] = params
    t4 = np.hstack([d["test"], 2 * np.ones((2000, 1))])
    s4 = np.hstack([d["test_sub"], 2 * np.ones((250, 1))])
    mu, var = w.predict(t4)
    ref = go.MFGP(g["X4"], g["y4"], params, F=3, gram=False)
    mu0, var0 = ref.predict(t4)
    v_, _, rho_, _ = go.split_mf_params(params, 3)
    scale = float(go.k_ar1_diag(t4[:1], v_, rho_)[0])
    assert mu.shape == (2000, 1) and normwise(mu, mu0) < TOL and normwise(var, var0, scale) < TOL
    cov = w.predict_covariance(s4)
    assert normwise(cov, ref.predict_covariance(s4), scale) < TOL
    assert cov.min() >= 1e-10                                 # emukit's element-wise clip
    if noise_mode == "mixed":
        assert measured("mf_vs_gram_golden_mean", normwise(mu, g["mu_mf"]), GRAM_TOL) < GRAM_TOL
        assert measured("mf_vs_gram_golden_var", normwise(var, g["var_mf"], scale), GRAM_TOL) < GRAM_TOL
        lo4 = np.hstack([d["test_sub"], np.zeros((250, 1))])
        a, b = w.predict(lo4)                                 # lowest fidelity: only k_0 and noise_0
        assert measured("mf0_vs_gram_golden_mean", normwise(a, g["mu_mf0"]), GRAM_TOL) < GRAM_TOL
        assert measured("mf0_vs_gram_golden_var", normwise(b, g["var_mf0"], scale), GRAM_TOL) < GRAM_TOL
    Kxx = m.kern.K(s4)
    v, ls, rho, _ = go.split_mf_params(params, 3)
    assert normwise(Kxx, go.k_ar1(s4, s4, v, ls, rho, gram=False, same=True)) < 1e-13
    c = w.copy()
    c.set_data(g["X4"][:50], g["y4"][:50, None])
    assert w.X.shape[0] == len(g["X4"]) and c.X.shape[0] == 50


def test_mfgp_two_fidelity_rho_and_matern(gpcore_mod, go):
    """BASELINE config 2 shape in miniature: F = 2, rho != 1, test grid at the top fidelity;
    and the Matern-3/2 AR1 variant the robot scripts use (...MFGP.py:656)."""
    L_ = gpcore_mod._lib
    rng = np.random.default_rng(11)
    X4, y = synth(rng, 600, 2)
    p = np.array([4.0, 2.0, 3.0, 2.5, 1.0, 1.5, 2.0, 2.0, 0.8, 0.05, 0.02])
    Xs4 = np.hstack([rng.uniform([0, 0, 0], [10, 20, 10], (777, 3)), np.ones((777, 1))])
    for kind, ok in ((L_.KIND_MF_AR1_RBF, go.KIND_RBF), (L_.KIND_MF_AR1_MAT32, go.KIND_MAT32)):
        core = gpcore_mod.GPCore(kind, 2, 0)
        core.set_hypers(p, 1e-8)
        core.set_data(X4, y)
        core.factor()
        mean, var = core.predict(Xs4, L_.INCLUDE_NOISE | L_.CLIP_DIAG)
        ref = go.MFGP(X4, y, p, F=2, kind=ok, gram=False)
        mu0, var0 = ref.predict(Xs4)
        assert normwise(mean, mu0[:, 0]) < TOL and normwise(var, var0[:, 0], 4.0) < TOL
        mo, _ = core.predict(Xs4, 0, want_var=False)
        assert normwise(mo, mean) < 1e-13     # mean-only and mean+var assemble K* in different kernels
        core.close()


def test_chunking_and_ragged_sizes(gpcore_mod, go):
    """M not a multiple of the tile or of the launch chunk; results must not depend on chunking."""
    L_ = gpcore_mod._lib
    rng = np.random.default_rng(13)
    X4, y = synth(rng, 333, 1)
    core = gpcore_mod.GPCore(L_.KIND_SF_RBF, 1, 0)
    core.set_hypers(SF_PARAMS, 1e-8)
    core.set_data(X4, y)
    core.factor()
    Xs4 = np.hstack([rng.uniform(0, 10, (1001, 3)), np.zeros((1001, 1))])
    m1, v1 = core.predict(Xs4, L_.INCLUDE_NOISE)
    core.set_chunk(256)
    m2, v2 = core.predict(Xs4, L_.INCLUDE_NOISE)
    assert np.array_equal(m1, m2) and np.array_equal(v1, v2)
    ref = go.SFGP(X4[:, :3], y, SF_PARAMS, gram=False)
    mu0, var0 = ref.predict(Xs4[:, :3], clip_diag=None)
    assert normwise(m1, mu0[:, 0]) < TOL and normwise(v1, var0[:, 0], 4.0) < TOL
    m0, v0 = core.predict(np.zeros((0, 4)), 0)
    assert m0.shape == (0,)
    one_m, one_v = core.predict(Xs4[:1], L_.INCLUDE_NOISE)
    assert one_m[0] == m1[0] and one_v[0] == v1[0]
    with pytest.raises(ValueError):
        core.set_chunk(100)
    core.close()


def test_posterior_properties_at_scale(gpcore_mod):
    """Size-independent properties at a BASELINE-like size (N = 2048, M = 20000, F = 2): the mean
    is linear in y; the variance does not depend on y; predicting AT the training inputs
    reproduces K alpha; 0 <= latent variance <= prior variance."""
    L_ = gpcore_mod._lib
    rng = np.random.default_rng(17)
    X4, y = synth(rng, 2048, 2)
    p = np.array([4.0, 2.0, 3.0, 2.5, 1.0, 1.5, 2.0, 2.0, 0.8, 0.05, 0.02])
    Xs4 = np.hstack([rng.uniform([0, 0, 0], [10, 20, 10], (20000, 3)), np.ones((20000, 1))])
    core = gpcore_mod.GPCore(L_.KIND_MF_AR1_RBF, 2, 0)
    core.set_hypers(p, 1e-8)
    y2 = rng.standard_normal(2048)
    out = []
    for yy in (y, y2, 2.0 * y - 3.0 * y2):
        core.set_data(X4, yy)
        core.factor()
        out.append(core.predict(Xs4, 0))
    assert normwise(out[2][0], 2.0 * out[0][0] - 3.0 * out[1][0]) < 1e-10
    assert np.array_equal(out[0][1], out[1][1])
    prior = np.where(Xs4[:, 3] == 1, 0.8 ** 2 * 4.0 + 1.0, 4.0)
    assert np.all(out[0][1] > -1e-9) and np.all(out[0][1] <= prior + 1e-9)
    core.set_data(X4, y)
    core.factor()
    mt, _ = core.predict(X4, 0)
    K = core.kernel_matrix(X4[:512], X4)
    assert normwise(mt[:512], K @ core.alpha()) < 1e-10
    core.close()


# ---------------------------------------------------------------------------------------------
# information gain
# ---------------------------------------------------------------------------------------------
def _cands_from_golden(g):
    rows, off = g["cand_rows"], g["cand_off"]
    return [rows[off[i]:off[i + 1]] for i in range(len(off) - 1)]


def test_ig_sequential_sf_golden(gpcore_mod, go):
    from gpcore.GPy.kern import RBF
    from gpcore.GPy.models import GPRegression
    from gpcore.infogain import seq_info_gain
    d, g = golden("field_data.npz"), golden("gp_oracle.npz")
    gp = GPRegression(d["Xh"], d["y"][:, None], RBF(3, ARD=True))
    gp.param_array[:] = SF_PARAMS
    cands = [c[:, :3] for c in _cands_from_golden(g)]
    I, best = seq_info_gain(gp, cands, SF_PARAMS[-1], first_preadded=True)
    assert measured("ig_sf_seq_vs_gram_golden", normwise(I, g["ig_sf_seq"]), GRAM_TOL) < GRAM_TOL   # literal refit loop, Gram-trick CPU
    ref = go.SFGP(d["Xh"], d["y"], SF_PARAMS, gram=False)
    I0 = np.array([go.ig_seq_sf_refit(ref, c, first_preadded=True) for c in cands[:4]])
    assert normwise(I[:4], I0) < TOL
    assert best == int(np.argmax(g["ig_sf_seq"]))
    I2, _ = seq_info_gain(gp, cands[:4], SF_PARAMS[-1], first_preadded=False)
    I20 = np.array([go.ig_seq_sf_refit(ref, c, first_preadded=False) for c in cands[:4]])
    assert normwise(I2, I20) < TOL


def test_ig_sequential_mf_golden(gpcore_mod, go):
    L_ = gpcore_mod._lib
    g = golden("gp_oracle.npz")
    core = gpcore_mod.GPCore(L_.KIND_MF_AR1_RBF, 3, 0)
    core.set_hypers(MF_PARAMS, 1e-8)
    core.set_data(g["X4"], g["y4"])
    core.factor()
    I, best = core.ig_seq(g["cand_rows"], g["cand_off"], MF_PARAMS[-1], pred_fid=0)
    assert measured("ig_mf_seq_vs_gram_golden", normwise(I, g["ig_mf_seq"]), GRAM_TOL) < GRAM_TOL
    ref = go.MFGP(g["X4"], g["y4"], MF_PARAMS, F=3, gram=False)
    cands = _cands_from_golden(g)
    I0 = np.array([go.ig_seq_mf_refit(ref, c, MF_PARAMS[-1], 0) for c in cands[:4]])
    assert normwise(I[:4], I0) < TOL
    assert best == int(np.argmax(g["ig_mf_seq"]))
    # queried at each point's own fidelity
    Iown, _ = core.ig_seq(g["cand_rows"], g["cand_off"], MF_PARAMS[-1], pred_fid=-1)
    c0 = cands[0]
    own = go.ig_seq_schur(ref, c0, ref.noise_of(c0), ref.noise_of(c0), MF_PARAMS[-1])
    assert abs(Iown[0] - own) < TOL * abs(own)
    core.close()


def test_ig_logdet_golden(gpcore_mod, go):
    L_ = gpcore_mod._lib
    d, g = golden("field_data.npz"), golden("gp_oracle.npz")
    cands = _cands_from_golden(g)
    # single fidelity
    core = gpcore_mod.GPCore(L_.KIND_SF_RBF, 1, 0)
    core.set_hypers(SF_PARAMS, 1e-8)
    core.set_data(np.hstack([d["Xh"], np.zeros((len(d["Xh"]), 1))]), d["y"])
    core.factor()
    grid4 = np.hstack([d["ig_grid"], np.zeros((len(d["ig_grid"]), 1))])
    rows = g["cand_rows"].copy(); rows[:, 3] = 0
    I, prior, best = core.ig_logdet(grid4, rows, g["cand_off"])
    assert measured("ig_sf_logdet_vs_gram_golden", normwise(I, g["ig_sf_ld"], 1.0), TOL) < TOL   # G = 300 log-dets of a Gram-trick refit
    ref = go.SFGP(d["Xh"], d["y"], SF_PARAMS, gram=False)
    I0 = np.array([go.ig_logdet_refit(ref, d["ig_grid"], c[:, :3]) for c in cands[:3]])
    assert measured("ig_sf_logdet_vs_refit", normwise(I[:3], I0, 1.0), TOL) < TOL
    _, cov = ref.predict(d["ig_grid"], full_cov=True)
    assert measured("ig_sf_logdet_prior", abs(prior - np.linalg.slogdet(cov)[1]) / abs(prior), TOL) < TOL
    assert best == int(np.argmax(g["ig_sf_ld"]))
    core.close()
    # multi fidelity, grid at fidelity 2
    core = gpcore_mod.GPCore(L_.KIND_MF_AR1_RBF, 3, 0)
    core.set_hypers(MF_PARAMS, 1e-8)
    core.set_data(g["X4"], g["y4"])
    core.factor()
    g4 = np.hstack([d["ig_grid"], 2 * np.ones((len(d["ig_grid"]), 1))])
    I, prior, best = core.ig_logdet(g4, g["cand_rows"], g["cand_off"])
    assert measured("ig_mf_logdet_vs_gram_golden", normwise(I, g["ig_mf_ld"], 1.0), TOL) < TOL
    assert best == int(np.argmax(g["ig_mf_ld"]))
    core.close()


def test_ig_edge_cases(gpcore_mod, go):
    """Empty candidate set, empty candidates, k = 1 and k = 64, > 64 rejected, ragged packing
    across tile and chunk boundaries, masks."""
    L_ = gpcore_mod._lib
    rng = np.random.default_rng(23)
    X4, y = synth(rng, 400, 3)
    core = gpcore_mod.GPCore(L_.KIND_MF_AR1_RBF, 3, 0)
    core.set_hypers(MF_PARAMS, 1e-8)
    core.set_data(X4, y)
    core.factor()
    ref = go.MFGP(X4, y, MF_PARAMS, F=3, gram=False)
    I, best = core.ig_seq(np.zeros((1, 4)), np.zeros(1, dtype=np.int64), 0.02)
    assert I.shape == (0,) and best == -1
    ks = [0, 1, 64, 3, 0, 50, 50, 50, 7, 64, 64, 2]
    cands = [np.hstack([rng.uniform([0, 0, 0], [10, 20, 10], (k, 3)), rng.integers(0, 3, (k, 1)).astype(float)]) for k in ks]
    rows, offs = gpcore_mod.GPCore._ragged(cands)
    I, best = core.ig_seq(rows, offs, 0.02, pred_fid=-1)
    assert I[0] == 0.0 and I[4] == 0.0
    for c in (1, 2, 3, 8):
        want = go.ig_seq_schur(ref, cands[c], ref.noise_of(cands[c]), ref.noise_of(cands[c]), 0.02)
        assert abs(I[c] - want) < TOL * max(1.0, abs(want)), c
    core.set_chunk(128)                                       # every tile its own chunk
    I2, best2 = core.ig_seq(rows, offs, 0.02, pred_fid=-1)
    assert np.array_equal(I, I2) and best == best2
    core.set_chunk(16384)
    with pytest.raises(ValueError):
        core.ig_seq(np.zeros((65, 4)), np.array([0, 65]), 0.02)
    # query at fidelity 0 needs 2k rows: k = 64 fits exactly one tile
    Ip, _ = core.ig_seq(rows, offs, 0.02, pred_fid=0)
    c = cands[2]; cp = c.copy(); cp[:, 3] = 0
    want = go.ig_seq_schur(ref, c, ref.noise_of(c), ref.noise_of(cp), 0.02, Xpred=cp)
    assert abs(Ip[2] - want) < TOL * abs(want)
    # mask: nothing conditioned on -> every point is scored against the data alone
    mask = np.zeros(rows.shape[0], dtype=np.uint8)
    Im, _ = core.ig_seq(rows, offs, 0.02, pred_fid=-1, row_mask=mask)
    _, v = ref.predict(cands[3])
    assert abs(Im[3] - np.sum(np.log(1 + v[:, 0] / 0.02))) < TOL * 10
    # log-det with an empty candidate list returns the prior only
    g4 = np.hstack([rng.uniform([0, 0, 0], [10, 20, 10], (50, 3)), 2 * np.ones((50, 1))])
    J, prior, b = core.ig_logdet(g4, np.zeros((1, 4)), np.zeros(1, dtype=np.int64))
    assert J.shape == (0,) and b == -1 and np.isfinite(prior)
    core.close()


def test_ig_agent_operators(gpcore_mod, go):
    """The CalcCost slot: reference method names on an agent-like object whose pathToTrajPoints is
    host code; batched score_many equals one-at-a-time scoring; guards of calcPathInfoSFBatch."""
    from gpcore.GPy.kern import RBF
    from gpcore.GPy.models import GPRegression
    from gpcore.emukit.multi_fidelity.kernels import LinearMultiFidelityKernel
    from gpcore.emukit.multi_fidelity.models import GPyLinearMultiFidelityModel
    from gpcore.emukit.model_wrappers.gpy_model_wrappers import GPyMultiOutputWrapper
    from gpcore.infogain import InfoGainOperators
    d, g = golden("field_data.npz"), golden("gp_oracle.npz")
    rng = np.random.default_rng(29)

    class Agent(InfoGainOperators):
        fidLevs = [0.25, 2.25, 6.25]

        def pathToTrajPoints(self, V, E, path, dense=False, t_off=0, withVar=False):
            p = E[path]          # the toy "edge store" is a dict path -> (k, 5) array x,y,z,t,var
            return p if withVar else p[:, :4]

    ag = Agent()
    ag.fieldGrid = d["ig_grid"]
    ag.sfgp = GPRegression(d["Xh"], d["y"][:, None], RBF(3, ARD=True))
    ag.sfgp.param_array[:] = SF_PARAMS
    k = LinearMultiFidelityKernel([RBF(3, ARD=True) for _ in range(3)])
    ag.mfgp = GPyMultiOutputWrapper(GPyLinearMultiFidelityModel(g["X4"], g["y4"][:, None], k, n_fidelities=3), 3, 1)
    ag.mfgp.gpy_model.param_array[:] = MF_PARAMS
    E = {}
    for c in range(6):
        kk = int(rng.integers(3, 12))
        a = rng.uniform([0, 0, 0], [10, 20, 10])
        pts = a[None] + np.linspace(0, 1, kk)[:, None] * rng.normal(0, 1.5, 3)[None]
        E[c] = np.hstack([pts, np.arange(kk)[:, None], rng.uniform(0, 7, (kk, 1))])
    paths = list(E)
    refsf = go.SFGP(d["Xh"], d["y"], SF_PARAMS, gram=False)
    refmf = go.MFGP(g["X4"], g["y4"], MF_PARAMS, F=3, gram=False)
    # sequential SF (calcPathInfoSF2)
    I, best = ag.score_many(None, E, paths, operator="calcPathInfoSF2")
    want = np.array([go.ig_seq_sf_refit(refsf, E[c][:, :3], first_preadded=True) for c in paths])
    assert normwise(I, want) < TOL and best == int(np.argmax(want))
    assert abs(ag.calcPathInfoSF2(None, E, 2) - want[2]) < TOL * abs(want[2])
    E[99] = E[0][:1]
    assert ag.calcPathInfoSF2(None, E, 99) == -np.inf
    # log-det SF (calcPathInfoSFBatch): max(., 0) and prior cache
    ag.logDetPrior = None
    J = ag.calcPathInfoSFBatch_many(None, E, paths)
    wantJ = np.array([max(go.ig_logdet_refit(refsf, d["ig_grid"], E[c][1:, :3]), 0) for c in paths])
    assert measured("agent_sfbatch_vs_refit", normwise(J, wantJ, 1.0), TOL) < TOL and ag.logDetPrior is not None
    # log-det MF (calculatePathInfoEmuBatch), fidelity labels from the variance column
    from gpcore.infogain import label_fidelity
    ag.logDetPrior = None
    Jm = ag.calculatePathInfoEmuBatch_many(None, E, paths)
    g4 = np.hstack([d["ig_grid"], 2 * np.ones((len(d["ig_grid"]), 1))])
    wantJm = np.array([go.ig_logdet_refit(refmf, g4, np.hstack([E[c][:, :3], label_fidelity(E[c][:, 4], ag.fidLevs)[:, None]]),
                                          clip_cov=1e-10)      # emukit predict_covariance clips element-wise
                       for c in paths])
    assert measured("agent_emubatch_clip_vs_refit", normwise(Jm, wantJm, 1.0), TOL) < TOL
    # sequential MF, un-windowed core (calculatePathInfoEmu with the window switched off)
    Is = ag.calculatePathInfoEmu_many(None, E, paths, windowed=False)
    wantIs = np.array([go.ig_seq_mf_refit(refmf, np.hstack([E[c][:, :3], label_fidelity(E[c][:, 4], ag.fidLevs)[:, None]]),
                                          MF_PARAMS[-1], 0) for c in paths])
    assert normwise(Is, wantIs) < TOL
    # windowed (literal reference semantics), checked against a literal loop on the oracle
    Iw = ag.calculatePathInfoEmu_many(None, E, paths[:3], windowed=True)
    lx, ly = MF_PARAMS[1], MF_PARAMS[2]
    for ci, c in enumerate(paths[:3]):
        X = np.hstack([E[c][:, :3], label_fidelity(E[c][:, 4], ag.fidLevs)[:, None]])
        allX = g["X4"].copy(); tot = 0.0
        for i in range(len(X)):
            allX = np.concatenate((allX, X[i:i + 1]))
            tempX = allX[np.logical_and(allX[:, 0] < 5 * lx, allX[:, 1] < 5 * ly)]
            tmp = go.MFGP(tempX, np.zeros(len(tempX)), MF_PARAMS, F=3, gram=False)
            q = X[i:i + 1].copy(); q[0, 3] = 0
            tot += np.log(1 + tmp.predict(q)[1][0, 0] / MF_PARAMS[-1])
        assert measured("agent_emu_windowed_vs_loop_%d" % ci, abs(Iw[ci] - tot) / abs(tot), TOL) < TOL, (ci, Iw[ci], tot)
    # reference quirk (PhysicalExperimentCode/GraceRIGV3.py:608-611): the single-path operator scores on a cached copy
    # mfgp2 whose DATA follow the agent but whose HYPER-PARAMETERS stay those of the moment the copy was made
    ag.logDetPrior = None
    first = ag.calculatePathInfoEmuBatch(None, E, paths[0])
    assert abs(first - Jm[0]) < 1e-9 * max(1.0, abs(Jm[0]))
    p_new = MF_PARAMS.copy()
    p_new[0] *= 1.7
    p_new[-1] *= 2.0
    ag.mfgp.gpy_model.param_array[:] = p_new
    stale = ag.calculatePathInfoEmuBatch(None, E, paths[1])
    assert abs(stale - Jm[1]) < 1e-9 * max(1.0, abs(Jm[1]))          # the copy still has the old hyper-parameters
    ag.reference_quirks = False
    ag.logDetPrior = None
    fresh = ag.calculatePathInfoEmuBatch(None, E, paths[1])
    assert abs(fresh - stale) > 1e-4                                  # the agent's current model differs


def test_hot_kernel_timing_hooks(gpcore_mod):
    L_ = gpcore_mod._lib
    rng = np.random.default_rng(31)
    X4, y = synth(rng, 512, 1)
    core = gpcore_mod.GPCore(L_.KIND_SF_RBF, 1, 0)
    core.set_hypers(SF_PARAMS, 1e-8)
    core.set_data(X4, y)
    core.factor()
    Xq = np.hstack([rng.uniform(0, 10, (4096, 3)), np.zeros((4096, 1))])
    core.enable_hot_timing(True)
    for mode, flops in ((L_.MODE_FP64, 4096 * 512 * (512 + 32)), (L_.MODE_INT8, 4096 * 512 * (512 + 64))):
        core.set_mode(mode)
        assert core.mode() == mode
        core.predict(Xq, 0)                             # first call of a mode may build its operand images
        core.hot_kernel_time(reset=True)
        n0 = core.launch_count()
        core.predict(Xq, 0)
        ms, n, fl = core.hot_kernel_time(reset=True)
        assert n == 1 and ms > 0 and fl == flops
        assert core.launch_count() - n0 == 3           # K* assembly, L^-1 K* contraction, finalize
    assert core.stream() is not None
    core.close()
