"""GPU parity of the tcgen05 / INT8 (Ozaki splitting) variance path against the FP64 DMMA path and
the oracle.  Same tolerance as everywhere: 1e-9 relative, normwise (max|d| <= 1e-9 max(|ref|, var))."""
import numpy as np
import pytest

from conftest import golden, measured, normwise

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


def synth(rng, N, F):
    X = rng.uniform([0, 0, 0], [10, 20, 10], (N, 3))
    f = rng.integers(0, F, (N, 1)).astype(float) if F > 1 else np.zeros((N, 1))
    y = np.sin(X[:, 0]) * np.cos(0.3 * X[:, 1]) + 0.2 * f[:, 0] + 0.05 * rng.standard_normal(N)
    return np.hstack([X, f]), y


@pytest.mark.parametrize("N,F,M", [(1, 1, 5), (100, 1, 333), (129, 3, 1000), (300, 1, 128), (709, 3, 2000),
                                     (1000, 2, 4097), (2048, 2, 20000)])
def test_int8_matches_fp64_and_oracle(gpcore_mod, go, N, F, M):
    L_ = gpcore_mod._lib
    rng = np.random.default_rng(1000 + N)
    X4, y = synth(rng, N, F)
    if F == 1:
        kind, p = L_.KIND_SF_RBF, SF_PARAMS
    elif F == 3:
        kind, p = L_.KIND_MF_AR1_RBF, MF_PARAMS
    else:
        kind, p = L_.KIND_MF_AR1_RBF, np.array([4.0, 2.0, 3.0, 2.5, 1.0, 1.5, 2.0, 2.0, 0.8, 0.05, 0.02])
    core = gpcore_mod.GPCore(kind, F, 0)
    core.set_hypers(p, 1e-8)
    core.set_data(X4, y)
    core.factor()
    Xs4 = np.hstack([rng.uniform([0, 0, 0], [10, 20, 10], (M, 3)), rng.integers(0, F, (M, 1)).astype(float)])
    Xs4[: min(M, N) // 2, :3] = X4[: min(M, N) // 2, :3] + 1e-3      # next to training points: var << prior (cancellation)
    flags = L_.INCLUDE_NOISE | L_.CLIP_DIAG
    core.set_mode(L_.MODE_FP64)
    m0, v0 = core.predict(Xs4, flags)
    core.set_mode(L_.MODE_INT8)
    m1, v1 = core.predict(Xs4, flags)
    scale = float(np.max(core.kernel_matrix(Xs4[:1], Xs4[:1]))) if F == 1 else 4.64
    assert normwise(m1, m0) < 1e-12
    assert normwise(v1, v0, scale) < 1e-10, normwise(v1, v0, scale)
    ref = go.SFGP(X4[:, :3], y, p, gram=False) if F == 1 else go.MFGP(X4, y, p, F=F, gram=False)
    mu, var = ref.predict(Xs4[:, :3] if F == 1 else Xs4)
    assert normwise(m1, mu[:, 0]) < TOL and normwise(v1, var[:, 0], scale) < TOL
    assert normwise(m0, mu[:, 0]) < TOL and normwise(v0, var[:, 0], scale) < TOL
    core.set_chunk(256)
    m2, v2 = core.predict(Xs4, flags)
    assert np.array_equal(m1, m2) and np.array_equal(v1, v2)        # independent of the launch chunking
    core.close()


def test_int8_matern_and_nigp_noisy_inputs(gpcore_mod, go):
    from gpcore.nigp import NIGP
    L_ = gpcore_mod._lib
    g, d = golden("nigp_field.npz"), golden("field_data.npz")
    sf = float(g["sigma_f"])
    m = NIGP(verbose=False)
    m.lengthscales_, m.sigma_f_, m.sigma_y_, m.sigma_x_ = g["ls"], sf, float(g["sigma_y"]), g["sigma_x"]
    m.X_train_, m.y_train_, m.noise_diag_train_ = d["Xh"], d["y"], g["noise_diag"]
    assert m._factor().mode() == L_.MODE_INT8                       # the default
    mean, var = m.predict(d["test"])
    assert normwise(mean, g["mean"]) < TOL and normwise(var, g["var"], sf) < TOL       # reference NIGP.py values
    _, var_in = m.predict(d["test"], Xs_input_noise=g["sigma_x"])
    assert normwise(var_in, g["var_in"], sf) < TOL
    rng = np.random.default_rng(77)
    X4, y = synth(rng, 500, 1)
    core = gpcore_mod.GPCore(L_.KIND_SF_MAT32, 1, 0)
    core.set_hypers(SF_PARAMS, 1e-8)
    core.set_data(X4, y)
    core.factor()
    Xs4 = np.hstack([rng.uniform(0, 10, (900, 3)), np.zeros((900, 1))])
    mu, var = go.SFGP(X4[:, :3], y, SF_PARAMS, kind=go.KIND_MAT32, gram=False).predict(Xs4[:, :3])
    m1, v1 = core.predict(Xs4, L_.INCLUDE_NOISE | L_.CLIP_DIAG)
    assert normwise(m1, mu[:, 0]) < TOL and normwise(v1, var[:, 0], 4.0) < TOL
    core.close()


def test_int8_extreme_scales(gpcore_mod, go):
    """Large and tiny kernel variances / noise: the fixed-point scales follow the hypers and the rows of L^-1."""
    L_ = gpcore_mod._lib
    rng = np.random.default_rng(5)
    X4, y = synth(rng, 400, 1)
    Xs4 = np.hstack([rng.uniform(0, 10, (700, 3)), np.zeros((700, 1))])
    for p in (np.array([2500.0, 2.0, 3.0, 2.5, 1e-4]), np.array([1e-3, 1.0, 1.0, 1.0, 1e-6]), np.array([4.0, 0.3, 0.3, 0.3, 2.0])):
        core = gpcore_mod.GPCore(L_.KIND_SF_RBF, 1, 0)
        core.set_hypers(p, 1e-8)
        core.set_data(X4, y * np.sqrt(p[0]))
        core.factor()
        ref = go.SFGP(X4[:, :3], y * np.sqrt(p[0]), p, gram=False)
        mu, var = ref.predict(Xs4[:, :3])
        m1, v1 = core.predict(Xs4, L_.INCLUDE_NOISE | L_.CLIP_DIAG)
        assert measured("extreme_scales_mean_%g" % p[0], normwise(m1, mu[:, 0]), TOL) < TOL, p
        assert measured("extreme_scales_var_%g" % p[0], normwise(v1, var[:, 0], p[0]), TOL) < TOL, p
        core.close()


@pytest.mark.parametrize("scale,noise", [(1.0, 1.0), (2500.0, 1e-3), (1e-3, 1.0)])
def test_int8_information_gain_matches_fp64(gpcore_mod, go, scale, noise):
    """The information-gain operators run end to end on the INT8 path (V emitted as a digit image, Gram and cross
    products from it); the FP64 DMMA path of the same handle is the cross-check, the oracle's literal refit loop the
    reference.  Kernel variances x scale and noise x noise exercise the fixed-point scale of V."""
    L_ = gpcore_mod._lib
    rng = np.random.default_rng(31)
    N, F = 600, 3
    X4, y = synth(rng, N, F)
    p = MF_PARAMS.copy()
    p[[0, 4, 8]] *= scale
    p[-3:] *= scale * noise
    core = gpcore_mod.GPCore(L_.KIND_MF_AR1_RBF, F, 0)
    core.set_hypers(p, 1e-8)
    core.set_data(X4, y * np.sqrt(scale))
    core.factor()
    ks = [1, 7, 32, 0, 64, 19, 33]
    cands = [np.hstack([rng.uniform([0, 0, 0], [10, 20, 10], (k, 3)), rng.integers(0, F, (k, 1)).astype(float)]) for k in ks]
    cands[1][:, :3] = X4[:7, :3] + 1e-3                    # next to training points: Schur complements ~ noise
    rows, offs = gpcore_mod.GPCore._ragged(cands)
    grid = rng.uniform([0, 0, 0], [10, 20, 10], (150, 3))
    g4 = np.hstack([grid, 2 * np.ones((150, 1))])
    out = {}
    for mode in (L_.MODE_FP64, L_.MODE_INT8):
        core.set_mode(mode)
        out[mode] = (core.ig_seq(rows, offs, float(p[-1]), pred_fid=0)[0], core.ig_logdet(g4, rows, offs)[0],
                     core.ig_logdet(g4, rows, offs, clip=True)[0], core.ig_selfgrid(rows, offs, pred_fid=2, clip=True)[0])
    # signal-to-noise 4e4 (scale 2500, noise 1e-3): cond(K) ~ 1e7; with the digit scales of round 2 the two paths stay
    # inside 1e-9 there as well (measured <= 5.3e-10)
    tol = TOL
    for nm, a, b in zip(("seq", "logdet", "logdet_clip", "selfgrid"), out[L_.MODE_FP64], out[L_.MODE_INT8]):
        assert measured("ig_int8_vs_fp64_%s_scale%g_noise%g" % (nm, scale, noise), normwise(b, a, 1.0), tol) < tol, (a, b)
    ref = go.MFGP(X4, y * np.sqrt(scale), p, F=F, gram=False)
    want = np.array([go.ig_logdet_refit(ref, g4, c) if len(c) else 0.0 for c in cands[:4]])
    assert measured("ig_int8_logdet_vs_refit_scale%g_noise%g" % (scale, noise), normwise(out[L_.MODE_INT8][1][:4], want, 1.0), TOL) < TOL
    wf = np.array(out[L_.MODE_FP64][1][:4])
    assert measured("ig_fp64_logdet_vs_refit_scale%g_noise%g" % (scale, noise), normwise(wf, want, 1.0), TOL) < TOL
    core.close()


def test_int8_size_limit_and_fallback(gpcore_mod):
    """N = 16384 is the largest training set the INT8 path takes (int32 accumulators: 6 digit pairs x K x 128^2 < 2^31);
    one row more and the same call runs on the FP64 DMMA path.  Both sides of the limit against each other on the same
    problem (size-independent property: the two contractions compute the same V), plus chunk-boundary row counts."""
    L_ = gpcore_mod._lib
    rng = np.random.default_rng(16384)
    X4, y = synth(rng, 16385, 1)
    Xs4 = np.hstack([rng.uniform(0, 10, (1025, 3)), np.zeros((1025, 1))])
    Xs4[:200, :3] = X4[:200, :3] + 1e-3
    flags = L_.INCLUDE_NOISE | L_.CLIP_DIAG
    res = {}
    for N in (16384, 16385):
        core = gpcore_mod.GPCore(L_.KIND_SF_RBF, 1, 0)
        core.set_hypers(SF_PARAMS, 1e-8)
        core.set_data(X4[:N], y[:N])
        core.factor()
        assert core.mode() == L_.MODE_INT8
        m1, v1 = core.predict(Xs4, flags)
        core.set_mode(L_.MODE_FP64)
        m0, v0 = core.predict(Xs4, flags)
        assert normwise(m1, m0) < 1e-12 and normwise(v1, v0, SF_PARAMS[0]) < 1e-9, N
        assert np.all(v1 >= SF_PARAMS[-1] * (1 - 1e-9)) and np.all(v1 <= (SF_PARAMS[0] + SF_PARAMS[-1]) * (1 + 1e-9))
        res[N] = (m1, v1)
        core.close()
    # one more training row barely moves the posterior at points far from it: the two problems agree loosely
    assert normwise(res[16384][0], res[16385][0]) < 1e-2


def test_host_predict_stages(gpcore_mod, go):
    """gpc_predict on host buffers longer than one device stage (2^17 rows): two alternating stages with the copies on
    a second stream, pageable caller arrays staged through the handle's pinned ring.  A ragged 600001-row call must
    equal the same rows predicted in small independent calls, and the same call on an input the caller pinned."""
    L_ = gpcore_mod._lib
    rng = np.random.default_rng(8)
    X4, y = synth(rng, 300, 1)
    core = gpcore_mod.GPCore(L_.KIND_SF_RBF, 1, 0)
    core.set_hypers(SF_PARAMS, 1e-8)
    core.set_data(X4, y)
    core.factor()
    M = 600001
    Xs4 = np.hstack([rng.uniform(0, 10, (M, 3)), np.zeros((M, 1))])
    flags = L_.INCLUDE_NOISE | L_.CLIP_DIAG
    m, v = core.predict(Xs4, flags)
    import torch
    Xp = torch.from_numpy(Xs4).pin_memory().numpy()                    # caller-pinned input: used in place
    mp, vp = core.predict(Xp, flags)
    assert np.array_equal(m, mp) and np.array_equal(v, vp)
    probes = [0, 1, 131071, 131072, 131073, 262143, 262144, 262145, 524287, 524288, 599999, 600000]
    for lo in (0, 130500, 262000, 524000, 599000):
        hi = min(M, lo + 1001)
        m0, v0 = core.predict(np.ascontiguousarray(Xs4[lo:hi]), flags)
        assert np.array_equal(m[lo:hi], m0) and np.array_equal(v[lo:hi], v0), lo
    assert np.all(np.isfinite(m[probes])) and np.all(v[probes] > 0)
    ref = go.SFGP(X4[:, :3], y, SF_PARAMS, gram=False)
    mu, var = ref.predict(Xs4[probes, :3])
    assert normwise(m[probes], mu[:, 0]) < TOL and normwise(v[probes], var[:, 0], SF_PARAMS[0]) < TOL
    core.close()


def test_fp32_tolerance_mode(gpcore_mod, go):
    """north_star's optional reduced-precision mode ("1e-4 for an optional FP32 mode"): GPC_MODE_INT8_F32 uses the four
    most significant digits of each operand (10 of the 21 digit GEMMs).  Posterior and information gain against the
    oracle at the stated 1e-4; the measured error (~1e-7: balanced digits truncate to nearest) is asserted two orders
    inside it.  The mean is formed in FP64 in every mode."""
    L_ = gpcore_mod._lib
    rng = np.random.default_rng(4242)
    N, F, M = 1500, 3, 6000
    X4, y = synth(rng, N, F)
    core = gpcore_mod.GPCore(L_.KIND_MF_AR1_RBF, F, 0)
    core.set_hypers(MF_PARAMS, 1e-8)
    core.set_data(X4, y)
    core.factor()
    Xs4 = np.hstack([rng.uniform([0, 0, 0], [10, 20, 10], (M, 3)), rng.integers(0, F, (M, 1)).astype(float)])
    Xs4[:700, :3] = X4[:700, :3] + 1e-3
    ref = go.MFGP(X4, y, MF_PARAMS, F=F, gram=False)
    mu, var = ref.predict(Xs4)
    flags = L_.INCLUDE_NOISE | L_.CLIP_DIAG
    core.set_mode(L_.MODE_INT8_F32)
    assert core.mode() == L_.MODE_INT8_F32
    m1, v1 = core.predict(Xs4, flags)
    em, ev = normwise(m1, mu[:, 0]), normwise(v1, var[:, 0], 4.64)
    print("INT8_F32 posterior: normwise mean %.2e var %.2e, element-wise var %.2e" % (em, ev, np.max(np.abs(v1 - var[:, 0]) / var[:, 0])))
    assert em < TOL                                  # FP64 mean
    assert ev < 1e-6 and np.max(np.abs(v1 - var[:, 0]) / var[:, 0]) < 1e-4
    cands = [np.hstack([rng.uniform([0, 0, 0], [10, 20, 10], (k, 3)), rng.integers(0, F, (k, 1)).astype(float)]) for k in (5, 32, 17, 9)]
    rows, offs = gpcore_mod.GPCore._ragged(cands)
    g4 = np.hstack([rng.uniform([0, 0, 0], [10, 20, 10], (120, 3)), 2 * np.ones((120, 1))])
    I, _ = core.ig_seq(rows, offs, float(MF_PARAMS[-1]), pred_fid=0)
    J, _, _ = core.ig_logdet(g4, rows, offs)
    I0 = np.array([go.ig_seq_mf_refit(ref, c, float(MF_PARAMS[-1]), 0) for c in cands])
    J0 = np.array([go.ig_logdet_refit(ref, g4, c) for c in cands])
    print("INT8_F32 information gain: seq %.2e log-det %.2e" % (normwise(I, I0), normwise(J, J0, 1.0)))
    assert normwise(I, I0) < 1e-4 and normwise(J, J0, 1.0) < 1e-4
    core.set_mode(L_.MODE_INT8)                      # and back: the full-precision path is untouched
    m2, v2 = core.predict(Xs4, flags)
    assert normwise(v2, var[:, 0], 4.64) < TOL
    core.close()


def test_host_predict_stages_with_per_row_input_noise(gpcore_mod):
    """NIGP.predict with a per-row Xs_input_noise (M, D) array on more rows than one device stage (2^17): the noise
    rows travel through the same pinned staging ring as the test rows.  A 300 001-row call must equal the same rows
    predicted in independent single-stage calls."""
    from gpcore.nigp import NIGP
    g, d = golden("nigp_field.npz"), golden("field_data.npz")
    m = NIGP(verbose=False)
    m.lengthscales_, m.sigma_f_, m.sigma_y_, m.sigma_x_ = g["ls"], float(g["sigma_f"]), float(g["sigma_y"]), g["sigma_x"]
    m.X_train_, m.y_train_, m.noise_diag_train_ = d["Xh"], d["y"], g["noise_diag"]
    rng = np.random.default_rng(12)
    M = 300001
    Xs = rng.uniform([0, 0, 0], [10, 20, 10], (M, 3))
    sx = rng.uniform(0.01, 0.3, (M, 3))
    mean, var = m.predict(Xs, Xs_input_noise=sx)
    assert np.all(np.isfinite(mean)) and np.all(var >= 1e-12)
    for lo in (0, 131000, 262000, 299000):
        hi = min(M, lo + 1101)
        m0, v0 = m.predict(np.ascontiguousarray(Xs[lo:hi]), Xs_input_noise=np.ascontiguousarray(sx[lo:hi]))
        assert np.array_equal(mean[lo:hi], m0) and np.array_equal(var[lo:hi], v0), lo
    # one shared (D,) noise vector gives the same as its tiled (M, D) form
    _, v1 = m.predict(Xs[:5000], Xs_input_noise=g["sigma_x"])
    _, v2 = m.predict(Xs[:5000], Xs_input_noise=np.tile(g["sigma_x"], (5000, 1)))
    assert np.allclose(v1, v2, rtol=1e-13, atol=0)


def test_five_level_mode_is_inside_the_normwise_tolerance_only(gpcore_mod):
    """GPC_MODE_INT8_L5 (15 of the 21 digit GEMMs): an opt-in speed / margin trade.  At configs[1]'s size its posterior
    variance stays inside the 1e-9 NORMWISE tolerance (measured ~5e-10) but not element-wise, which is why the default
    keeps all six levels."""
    import scale_cases as sc
    L_ = gpcore_mod._lib
    X4, y, p, Xs4 = sc.configs1_inputs()
    o, _ = sc.frozen_or_live("c1", sc.configs1_oracle, sc.sha(X4, y, p, Xs4))
    core = gpcore_mod.GPCore(L_.KIND_MF_AR1_RBF, 2, 0)
    core.set_hypers(p, 1e-8)
    core.set_data(X4, y)
    core.factor()
    core.set_mode(L_.MODE_INT8_L5)
    m1, v1 = core.predict(Xs4, L_.INCLUDE_NOISE | L_.CLIP_DIAG)
    scale = float(p[0] * p[8] ** 2 + p[4])
    nm, nv = normwise(m1, o["c1_mu"]), normwise(v1, o["c1_var"], scale)
    print("INT8_L5: normwise mean %.2e var %.2e, element-wise var %.2e" % (nm, nv, np.max(np.abs(v1 - o["c1_var"]) / o["c1_var"])))
    assert nm < TOL and nv < TOL
    core.close()
