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
