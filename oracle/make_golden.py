"""Generate the committed golden fixtures under ``tests/golden/``.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs ``/root/reference``):

    python oracle/make_golden.py

* ``nigp_demo.npz``  -- the reference ``NIGP.py`` module, imported VERBATIM through
  ``oracle/gpy_shim``, run on its own ``__main__`` data (seed 0, 1-D sin, N = 40,
  ``NIGP.py:339-352``): fitted hypers, per-point noise, predictions (mean / var / cov, with and
  without test-input noise), NLML and posterior-mean gradients at the fitted hypers.
* ``nigp_field.npz`` -- the same reference functions on a bundled 3-D dataset
  (``GPData_0.2_fieldMeas_0_T0_0.csv``, 709 rows, estimated positions) with hypers held fixed.
* ``field_data.npz`` -- that dataset's columns (inputs for the SF / MF parity tests) and the
  reference's test grids (``exploreSimSettings.py:116-119``, ``exploreExpSettings.py:164-167``).
* ``traj_paths.npz`` -- the reference's own ``GraceAgent.pathToTrajPoints`` / ``evaluateTraj``
  (``GraceRIGV3.py:235-294,373-427``, imported through ``oracle/mpl_shim``) on seeded primitive chains.
* ``ig_operators.npz`` -- the reference's own path-cost operators (root and PhysicalExperimentCode
  ``GraceRIGV3.py``), run unmodified over adapters of the restated GP models.
* ``eid.npz`` -- the reference's own ``exploreSimSettings.getEID`` over adapters of the restated GP models.
* ``gp_oracle.npz``  -- outputs of the NumPy restatement (``gp_oracle.py``; GPy / emukit arithmetic,
  PARITY UNPINNED) on the same data: SF / MF predictions, covariances, information gains.  These
  freeze the restatement so a later edit to the oracle cannot silently move the target.
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")


def reference_nigp():
    sys.path.insert(0, os.path.join(HERE, "gpy_shim"))
    sys.path.insert(0, REF)
    import NIGP as ref  # the reference module itself
    assert os.path.realpath(ref.__file__).startswith(REF), ref.__file__
    return ref


def load_field(name="GPData_0.2_fieldMeas_0_T0_0.csv"):
    path = os.path.join(REF, "Data", "TrajectoriesAndEstimates", "GPDataSets", name)
    with open(path) as f:
        hdr = f.readline().strip().split(",")
        d = np.loadtxt(f, delimiter=",")
    col = lambda *n: d[:, [hdr.index(k) for k in n]]
    return {"t": col("t")[:, 0], "X": col("x", "y", "z"), "Xh": col("xh", "yh", "zh"),
            "y": col("fieldVal")[:, 0], "fidLev": col("fidLev")[:, 0].astype(int)}


def grid(specs):
    g = np.meshgrid(*[np.linspace(a, b, n) for a, b, n in specs])
    return np.array([gi.ravel("F") for gi in g]).T


def traj_golden():
    """``traj_paths.npz``: outputs of the reference's OWN ``GraceAgent.pathToTrajPoints`` (imported from
    /root/reference/GraceRIGV3.py through an empty matplotlib stand-in, ``oracle/mpl_shim``) on seeded
    random primitive chains, for dense / way-point mode with and without the variance column."""
    import types
    sys.path.insert(0, os.path.join(HERE, "mpl_shim"))
    sys.path.insert(0, os.path.join(HERE, "gpy_shim"))
    sys.path.insert(0, REF)
    import GraceRIGV3 as IG
    assert os.path.realpath(IG.__file__).startswith(REF)
    ag = IG.GraceAgent()
    ag.varianceRate, ag.measRate = 0.01, 0.2
    names = ag.legTypes
    rng = np.random.default_rng(2024)
    V, E, paths = [], {}, []
    for c in range(24):
        ne = int(rng.integers(1, 4))
        pos = rng.uniform([0, 0], [10, 20])
        first = len(V)
        V.append(types.SimpleNamespace(state=np.array([[pos[0]], [pos[1]]])))
        path = []
        for e in range(ne):
            pos = pos + rng.normal(0, 2.0, 2)
            V.append(types.SimpleNamespace(state=np.array([[pos[0]], [pos[1]]])))
            prims = []
            for _ in range(int(rng.integers(1, 6))):
                kind = int(rng.integers(0, 4))
                if kind == 0:
                    prims.append((names[0], float(rng.uniform(-1, 2)), 0.0, float(rng.uniform(0.05, 0.3))))
                elif kind == 1:
                    prims.append((names[1], float(rng.uniform(0.3, 1.0)), float(rng.uniform(-1.5, 2)), float(rng.uniform(0.05, 0.3))))
                elif kind == 2:
                    prims.append((names[2], float(rng.uniform(0.2, 3.0)), float(rng.uniform(0.1, 0.5))))
                else:
                    prims.append((names[3], float(rng.uniform(-1, 2)), float(rng.uniform(0.05, 0.3))))
            i1, i2 = first + e, first + e + 1
            E[(i1, i2)] = [(i1, i2, 0.0, 0.0, 0.0, 0.0, prims)]
            path.append((i1, i2, 0))
        paths.append(path)
    out = {}
    edge_off, xy, prim_off, pr = [0], [], [0], []
    for path in paths:
        edge_off.append(edge_off[-1] + len(path))
        for i1, i2, k in path:
            xy.append([V[i1].state[0, 0], V[i1].state[1, 0], V[i2].state[0, 0], V[i2].state[1, 0]])
            for q in E[(i1, i2)][k][-1]:
                vals = [float(names.index(q[0]))] + [float(v) for v in q[1:]]
                pr.append(vals + [0.0] * (4 - len(vals)))
            prim_off.append(len(pr))
    out.update(edge_off=np.array(edge_off), edge_xy=np.array(xy), prim_off=np.array(prim_off), prims=np.array(pr),
               variance_rate=ag.varianceRate, meas_rate=ag.measRate)
    for dense in (0, 1):
        for wv in (0, 1):
            rows, offs = [], [0]
            for path in paths:
                p = ag.pathToTrajPoints(V, E, path, dense=bool(dense), withVar=bool(wv))
                rows.append(p)
                offs.append(offs[-1] + len(p))
            out["pts_d%d_v%d" % (dense, wv)] = np.concatenate(rows)
            out["off_d%d_v%d" % (dense, wv)] = np.array(offs)
    np.savez_compressed(os.path.join(OUT, "traj_paths.npz"), **out)
    print("traj_paths:", {k: v.shape for k, v in out.items() if hasattr(v, "shape") and v.ndim})


def ig_operator_golden():
    """``ig_operators.npz``: the reference's OWN path-cost operators (``GraceRIGV3.py:443-562`` and
    ``PhysicalExperimentCode/GraceRIGV3.py:446-678``), imported and run unmodified, with the GP model
    objects they call replaced by thin GPy / emukit-shaped adapters over ``gp_oracle`` (GPy and emukit are
    not installable).  This pins the OPERATOR logic -- loops, pre-appends, windows, fidelity labels, guards --
    to the reference's code; the GP arithmetic underneath stays the (unpinned) restatement."""
    import copy
    import importlib.util
    import types
    sys.path.insert(0, os.path.join(HERE, "mpl_shim"))
    sys.path.insert(0, os.path.join(HERE, "gpy_shim"))
    sys.path.insert(0, REF)
    sys.path.insert(0, ROOT)
    import GraceRIGV3 as IGroot
    from oracle import gp_oracle as go
    spec = importlib.util.spec_from_file_location("GraceRIGV3_phys", os.path.join(REF, "PhysicalExperimentCode", "GraceRIGV3.py"))
    sys.path.insert(0, os.path.join(REF, "PhysicalExperimentCode"))
    IGphys = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(IGphys)

    class SF:
        def __init__(self, X, Y, params):
            self.params = np.asarray(params, float)
            self.Gaussian_noise = types.SimpleNamespace(variance=np.array([self.params[-1]]))
            self.kern = types.SimpleNamespace(lengthscale=self.params[1:4].copy(), variance=np.array([self.params[0]]))
            self.set_XY(X, Y)

        def set_XY(self, X, Y):
            self.X, self.Y = np.asarray(X, float), np.asarray(Y, float).reshape(-1, 1)
            self._gp = go.SFGP(self.X, self.Y, self.params, gram=False)

        def predict(self, Xn, full_cov=False):
            return self._gp.predict(np.asarray(Xn, float), full_cov=bool(full_cov))

        def copy(self):
            return copy.deepcopy(self)

    class MF:
        def __init__(self, X4, Y, params):
            self.params = np.asarray(params, float)
            v, ls, rho, _ = go.split_mf_params(self.params, 3)
            kern = types.SimpleNamespace(rbf=types.SimpleNamespace(lengthscale=ls[0].copy()),
                                         K=lambda A: go.k_ar1(A, A, v, ls, rho, gram=False, same=True))
            self.gpy_model = types.SimpleNamespace(param_array=self.params, kern=kern)
            self.set_data(X4, Y)

        def set_data(self, X4, Y):
            self.X, self.Y = np.asarray(X4, float), np.asarray(Y, float).reshape(-1, 1)
            self._gp = go.MFGP(self.X, self.Y, self.params, F=3, gram=False)

        def predict(self, X4):
            return self._gp.predict(np.asarray(X4, float))

        def predict_covariance(self, X4):
            return self._gp.predict_covariance(np.asarray(X4, float))

        def copy(self):
            return copy.deepcopy(self)

    fld = load_field()
    sel = np.r_[0:60, 300:360, 600:660]                       # 180 points, all three fidelity levels
    Xh, yv, lev = fld["Xh"][sel], fld["y"][sel], fld["fidLev"][sel]
    sf_params = np.array([4.0, 2.0, 3.0, 2.5, 0.05])
    mf_params = np.array([3.0, 2.5, 3.5, 3.0, 1.0, 1.5, 2.0, 2.0, 0.5, 1.0, 1.5, 1.5, 0.9, 1.1, 0.08, 0.04, 0.02])
    X4 = np.hstack([Xh, (3 - lev)[:, None].astype(float)])     # fidLev 3/2/1 -> index 0/1/2 (GPTrainers.py:55-61)
    g = np.load(os.path.join(OUT, "traj_paths.npz"))
    eo, po = g["edge_off"], g["prim_off"]
    names = ["Spiral", "Glide", "Swim", "FlatDive"]
    V, E, paths = [], {}, []
    for c in range(10):
        path = []
        for e in range(eo[c], eo[c + 1]):
            xy = g["edge_xy"][e]
            i1 = len(V); V.append(types.SimpleNamespace(state=np.array([[xy[0]], [xy[1]]])))
            i2 = len(V); V.append(types.SimpleNamespace(state=np.array([[xy[2]], [xy[3]]])))
            prims = []
            for q in g["prims"][po[e]:po[e + 1]]:
                k = int(q[0])
                prims.append((names[k],) + tuple(float(v) for v in (q[1:4] if k < 2 else q[1:3])))
            E[(i1, i2)] = [(i1, i2, 0.0, 0.0, 0.0, 0.0, prims)]
            path.append((i1, i2, 0))
        paths.append(path)
    grid_ig = grid([[0, 10, 5], [0, 20, 4], [0, 10, 3]])       # a small field grid (60 points)
    fid_levs = [0.05, 0.15, 0.3]
    out = dict(Xh=Xh, y=yv, X4=X4, sf_params=sf_params, mf_params=mf_params, grid=grid_ig, fidLevs=np.array(fid_levs),
               n_paths=len(paths), variance_rate=0.01, meas_rate=0.2)

    def agent_for(mod):
        ag = mod.GraceAgent()
        ag.varianceRate, ag.measRate, ag.fidLevs, ag.fieldGrid = 0.01, 0.2, fid_levs, grid_ig
        ag.sfgp, ag.mfgp = SF(Xh, yv, sf_params), MF(X4, yv, mf_params)
        return ag

    root, phys = agent_for(IGroot), agent_for(IGphys)
    ops = {"root_calcPathInfoSF2": lambda p: root.calcPathInfoSF2(V, E, p),
           "root_calcPathInfoSF": lambda p: root.calcPathInfoSF(V, E, p),
           "root_calculatePathInfoEmu": lambda p: root.calculatePathInfoEmu(V, E, p),
           "root_calculatePathInfoEmu2": lambda p: root.calculatePathInfoEmu2(V, E, p),
           "phys_calcPathInfoSF4": lambda p: phys.calcPathInfoSF4(V, E, p),
           "phys_calculatePathInfoEmu": lambda p: phys.calculatePathInfoEmu(V, E, p)}
    for name, fn in ops.items():
        vals = []
        for p in paths:
            for ag in (root, phys):                          # every call starts from the agent's own data
                ag.sfgp, ag.mfgp = SF(Xh, yv, sf_params), MF(X4, yv, mf_params)
                if hasattr(ag, "sfgp2"):
                    ag.sfgp2 = None
            vals.append(float(fn(p)))
        out[name] = np.array(vals)
        print(name, np.round(out[name][:4], 6))
    for name in ("calcPathInfoSFBatch", "calculatePathInfoEmuBatch"):
        vals = []
        for p in paths:
            phys.sfgp, phys.mfgp = SF(Xh, yv, sf_params), MF(X4, yv, mf_params)
            phys.sfgp2 = None
            phys.mfgp2 = None
            phys.logDetPrior = None
            vals.append(float(getattr(phys, name)(V, E, p)))
        out["phys_" + name] = np.array(vals)
        print(name, np.round(out["phys_" + name][:4], 6))
    # ---- calcPathInfoSFBatch as the planner drives it: logDetPrior reset once, then consecutive calls --
    # the cached copy keeps the points of the earlier calls (PhysicalExperimentCode/GraceRIGV3.py:590)
    phys.sfgp, phys.mfgp = SF(Xh, yv, sf_params), MF(X4, yv, mf_params)
    phys.sfgp2, phys.logDetPrior = None, None
    out["phys_calcPathInfoSFBatch_consecutive"] = np.array([float(phys.calcPathInfoSFBatch(V, E, p)) for p in paths])
    print("consecutive", np.round(out["phys_calcPathInfoSFBatch_consecutive"][:4], 6))
    # ---- small training sets: the sets of the windowed operators cross the 100-row threshold INSIDE a path
    # (N = 60), and a training set with no row inside the x < 3 lx, y < 3 ly window (empty-window fallbacks)
    def variant(tag, idx, names_):
        Xs_, ys_, X4s_ = Xh[idx], yv[idx], X4[idx]
        out[tag + "_idx"] = np.asarray(idx)
        r_, p_ = IGroot.GraceAgent(), IGphys.GraceAgent()
        for ag in (r_, p_):
            ag.varianceRate, ag.measRate, ag.fidLevs, ag.fieldGrid = 0.01, 0.2, fid_levs, grid_ig
        table = {"root_calcPathInfoSF": (r_, "calcPathInfoSF"), "phys_calcPathInfoSF4": (p_, "calcPathInfoSF4"),
                 "root_calculatePathInfoEmu": (r_, "calculatePathInfoEmu"), "phys_calculatePathInfoEmu": (p_, "calculatePathInfoEmu")}
        for nm in names_:
            ag, meth = table[nm]
            vals = []
            for pth in paths:
                ag.sfgp, ag.mfgp = SF(Xs_, ys_, sf_params), MF(X4s_, ys_, mf_params)
                if hasattr(ag, "sfgp2"):
                    ag.sfgp2 = None
                vals.append(float(getattr(ag, meth)(V, E, pth)))
            out[tag + "_" + nm] = np.array(vals)
            print(tag, nm, np.round(out[tag + "_" + nm][:4], 6))

    variant("n60", np.r_[0:20, 60:80, 120:140],
            ["root_calcPathInfoSF", "phys_calcPathInfoSF4", "root_calculatePathInfoEmu", "phys_calculatePathInfoEmu"])
    variant("n97", np.r_[0:33, 60:92, 120:152], ["root_calcPathInfoSF", "phys_calcPathInfoSF4", "root_calculatePathInfoEmu"])
    far = np.where(np.logical_or(Xh[:, 0] >= 3 * sf_params[1], Xh[:, 1] >= 3 * sf_params[2]))[0]
    variant("nowin", far[:120], ["root_calcPathInfoSF", "phys_calcPathInfoSF4"])
    variant("nowin40", far[:40], ["root_calcPathInfoSF", "phys_calcPathInfoSF4"])
    np.savez_compressed(os.path.join(OUT, "ig_operators.npz"), **out)


def eid_golden():
    """``eid.npz``: the reference's own ``exploreSimSettings.getEID`` (imported from a scratch working directory --
    the module writes ``Data/fieldSettings.txt`` at import) over adapters of the restated GP models."""
    import tempfile
    import types
    sys.path.insert(0, os.path.join(HERE, "mpl_shim"))
    sys.path.insert(0, os.path.join(HERE, "gpy_shim"))
    sys.path.insert(0, REF)
    sys.path.insert(0, ROOT)
    from oracle import gp_oracle as go
    cwd = os.getcwd()
    tmp = tempfile.mkdtemp()
    os.makedirs(os.path.join(tmp, "Data"))
    os.chdir(tmp)
    try:
        import exploreSimSettings as ess
    finally:
        os.chdir(cwd)
    fld = load_field()
    sel = np.r_[0:60, 300:360, 600:660]
    Xh, yv, lev = fld["Xh"][sel], fld["y"][sel], fld["fidLev"][sel]
    sf_params = np.array([4.0, 2.0, 3.0, 2.5, 0.05])
    mf_params = np.array([3.0, 2.5, 3.5, 3.0, 1.0, 1.5, 2.0, 2.0, 0.5, 1.0, 1.5, 1.5, 0.9, 1.1, 0.08, 0.04, 0.02])
    X4 = np.hstack([Xh, (3 - lev)[:, None].astype(float)])
    sf = go.SFGP(Xh, yv[:, None], sf_params, gram=False)
    mf = go.MFGP(X4, yv[:, None], mf_params, F=3, gram=False)
    sfa = types.SimpleNamespace(predict=lambda X: sf.predict(np.asarray(X, float)),
                                kern=types.SimpleNamespace(variance=np.array([sf_params[0]])),
                                Gaussian_noise=types.SimpleNamespace(variance=np.array([sf_params[-1]])))
    mfa = types.SimpleNamespace(predict=lambda X: mf.predict(np.asarray(X, float)),
                                gpy_model=types.SimpleNamespace(param_array=mf_params))
    WS, mD = np.array([[0, 10], [0, 20]]), 10
    out = dict(Xh=Xh, y=yv, X4=X4, sf_params=sf_params, mf_params=mf_params, WS=WS, mD=mD)
    for auto in (0, 1):
        ess.auto = auto
        out["sf_auto%d" % auto], out["grid"] = ess.getEID(sfa, WS, mD)
        out["mf_auto%d" % auto], _ = ess.getEID(mfa, WS, mD, emu=True)
    ess.auto = 0
    out["sf_alpha03"], _ = ess.getEID(sfa, WS, mD, alpha=0.3)   # (an ndarray testSet raises in the reference: `testSet==None`)
    np.savez_compressed(os.path.join(OUT, "eid.npz"), **out)
    print("eid:", {k: v.shape for k, v in out.items() if hasattr(v, "shape") and v.ndim})


def main():
    os.makedirs(OUT, exist_ok=True)
    warnings.simplefilter("ignore")
    if "--traj-only" in sys.argv:
        traj_golden()
        return
    if "--eid-only" in sys.argv:
        eid_golden()
        return
    if "--ig-only" in sys.argv:
        ig_operator_golden()
        return
    traj_golden()
    ig_operator_golden()
    eid_golden()
    ref = reference_nigp()

    # ---- (i) NIGP demo, seed 0 -------------------------------------------------------------
    np.random.seed(0)
    N = 40
    X_true = np.linspace(-3, 3, N)[:, None]
    f_true = np.sin(X_true).ravel()
    sigma_x_true, sigma_y_true = 0.2, 0.05
    X_noisy = X_true + sigma_x_true * np.random.randn(*X_true.shape)
    y = f_true + sigma_y_true * np.random.randn(N)
    m = ref.NIGP(n_restarts=2, iters=10, verbose=False)
    m.fit(X_noisy, y)
    Xs = np.linspace(-4, 4, 200)[:, None]
    mean, var = m.predict(Xs)
    _, cov = m.predict(Xs, return_cov=True)
    _, var_in = m.predict(Xs, Xs_input_noise=m.sigma_x_)
    _, cov_in = m.predict(Xs[:50], Xs_input_noise=np.full((50, 1), 0.1), return_cov=True)
    fm, grads = ref.compute_post_mean_and_gradients(X_noisy, y, m.lengthscales_, m.sigma_f_, m.sigma_y_,
                                                    noise_diag=m.noise_diag_train_)
    log_hyp = np.log(np.concatenate([m.lengthscales_, [m.sigma_f_, m.sigma_y_], m.sigma_x_]))
    nlml = ref.neg_log_marginal_likelihood(log_hyp, X_noisy, y, grads)
    K = ref.SE_ARD_kernel(X_noisy, Xs, m.lengthscales_, m.sigma_f_)
    np.savez_compressed(os.path.join(OUT, "nigp_demo.npz"), X=X_noisy, y=y, Xs=Xs,
                        lengthscales=m.lengthscales_, sigma_f=m.sigma_f_, sigma_y=m.sigma_y_,
                        sigma_x=m.sigma_x_, noise_diag=m.noise_diag_train_, params=m.get_params(),
                        mean=mean, var=var, cov=cov, var_in=var_in, cov_in=cov_in,
                        f_mean_train=fm, grads=grads, log_hyp=log_hyp, nlml=nlml, K=K)
    print("nigp_demo: hypers", m.get_params())

    # ---- (ii) NIGP on the bundled 3-D dataset, fixed hypers ------------------------------------
    fld = load_field()
    Xh, yv = fld["Xh"], fld["y"]
    ls = np.array([2.0, 3.0, 2.5]); sf = 4.0; sy = 0.2; sx = np.array([0.1, 0.1, 0.05])
    fm0, g0 = ref.compute_post_mean_and_gradients(Xh, yv, ls, sf, sy, noise_diag=None)
    nd = np.sum(g0 ** 2 * sx[None, :] ** 2, axis=1)
    lh = np.log(np.concatenate([ls, [sf, sy], sx]))
    nlml_f = ref.neg_log_marginal_likelihood(lh, Xh, yv, g0)
    nlml_f_extra = ref.neg_log_marginal_likelihood(lh, Xh, yv, g0, 0.01 * np.ones(len(yv)))
    mm = ref.NIGP(verbose=False)
    mm.lengthscales_, mm.sigma_f_, mm.sigma_y_, mm.sigma_x_ = ls, sf, sy, sx
    mm.X_train_, mm.y_train_, mm.noise_diag_train_ = Xh, yv, nd
    test = grid([[0, 10, 10], [0, 20, 20], [0, 10, 10]])     # exploreSimSettings.py:116-119
    sub = test[::8]                                          # 250 points for the full covariance
    mean_f, var_f = mm.predict(test)
    _, cov_f = mm.predict(sub, return_cov=True)
    _, var_f_in = mm.predict(test, Xs_input_noise=sx)
    mean_only = mm.predict(sub, return_var=False)
    np.savez_compressed(os.path.join(OUT, "nigp_field.npz"), ls=ls, sigma_f=sf, sigma_y=sy, sigma_x=sx,
                        f_mean_train=fm0, grads=g0, noise_diag=nd, log_hyp=lh, nlml=nlml_f,
                        nlml_extra=nlml_f_extra, mean=mean_f, var=var_f, cov_sub=cov_f, var_in=var_f_in,
                        mean_only_sub=mean_only)
    print("nigp_field: nlml", nlml_f)

    # ---- (iii) dataset + grids ---------------------------------------------------------------
    ig_grid = grid([[0, 10, 10], [0, 20, 6], [0, 10, 5]])    # exploreExpSettings.py:164-167 scaled to the box
    np.savez_compressed(os.path.join(OUT, "field_data.npz"), test=test, test_sub=sub, ig_grid=ig_grid, **fld)

    # ---- (iv) the NumPy restatement on the same data (freeze) ------------------------------------
    sys.path.insert(0, ROOT)
    from oracle import gp_oracle as go
    sf_params = np.array([4.0, 2.0, 3.0, 2.5, 0.05])
    gp = go.SFGP(Xh, yv, sf_params)
    mu_sf, var_sf = gp.predict(test)
    _, cov_sf = gp.predict(sub, full_cov=True)
    gp32 = go.SFGP(Xh, yv, sf_params, kind=go.KIND_MAT32)
    mu_32, var_32 = gp32.predict(test)
    # three fidelities as GPTrainers.py:55-61: index 0 = fidLev 3 (lowest) ... index 2 = fidLev 1
    order = [3, 2, 1]
    X4 = np.concatenate([np.hstack([Xh[fld["fidLev"] == lv], np.full((np.sum(fld["fidLev"] == lv), 1), float(i))])
                         for i, lv in enumerate(order)])
    y4 = np.concatenate([yv[fld["fidLev"] == lv] for lv in order])
    mf_params = np.array([3.0, 2.5, 3.5, 3.0, 1.0, 1.5, 2.0, 2.0, 0.5, 1.0, 1.5, 1.5, 0.9, 1.1, 0.08, 0.04, 0.02])
    mf = go.MFGP(X4, y4, mf_params, F=3)
    t4 = np.hstack([test, 2 * np.ones((len(test), 1))])
    s4 = np.hstack([sub, 2 * np.ones((len(sub), 1))])
    mu_mf, var_mf = mf.predict(t4)
    cov_mf = mf.predict_covariance(s4)
    mu_mf0, var_mf0 = mf.predict(np.hstack([sub, np.zeros((len(sub), 1))]))
    rng = np.random.default_rng(7)
    cands, cands4 = [], []
    for c in range(12):
        k = int(rng.integers(2, 20))
        a = rng.uniform([0, 0, 0], [10, 20, 10]); b = a + rng.normal(0, 1.0, 3)
        pts = a[None] + np.linspace(0, 1, k)[:, None] * (b - a)[None]
        cands.append(pts)
        cands4.append(np.hstack([pts, rng.integers(0, 3, (k, 1)).astype(float)]))
    ig_sf_seq = np.array([go.ig_seq_sf_refit(gp, c, first_preadded=True) for c in cands])
    ig_mf_seq = np.array([go.ig_seq_mf_refit(mf, c, sig_n=mf_params[-1], pred_fid=0) for c in cands4])
    ig_sf_ld = np.array([go.ig_logdet_refit(gp, ig_grid, c) for c in cands])
    g4 = np.hstack([ig_grid, 2 * np.ones((len(ig_grid), 1))])
    ig_mf_ld = np.array([go.ig_logdet_refit(mf, g4, c) for c in cands4])
    cand_rows = np.concatenate(cands4)
    cand_off = np.cumsum([0] + [len(c) for c in cands4])
    np.savez_compressed(os.path.join(OUT, "gp_oracle.npz"), sf_params=sf_params, mf_params=mf_params, X4=X4, y4=y4,
                        mu_sf=mu_sf, var_sf=var_sf, cov_sf=cov_sf, mu_32=mu_32, var_32=var_32,
                        mu_mf=mu_mf, var_mf=var_mf, cov_mf=cov_mf, mu_mf0=mu_mf0, var_mf0=var_mf0,
                        nlml_sf=gp.f.nlml, nlml_mf=mf.f.nlml, cand_rows=cand_rows, cand_off=cand_off,
                        ig_sf_seq=ig_sf_seq, ig_mf_seq=ig_mf_seq, ig_sf_ld=ig_sf_ld, ig_mf_ld=ig_mf_ld)
    print("gp_oracle: ig_sf_seq", ig_sf_seq[:3], "ig_mf_ld", ig_mf_ld[:3])
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
