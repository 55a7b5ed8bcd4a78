"""Information-gain path-cost operators (the ``agent.CalcCost`` slot of the RIG planner,
``GraceRIGV3.py:24,1099,1158``), batched on the GPU.

Two layers:

* array level -- ``seq_info_gain`` / ``logdet_info_gain`` score MANY candidates (lists of
  point arrays) in one call through ``gpc_ig_seq`` / ``gpc_ig_logdet``;
* agent level -- ``InfoGainOperators`` is a mixin with the reference's method names and
  ``(V, E, path, dense) -> float`` signatures (``calcPathInfoSF2``, ``calcPathInfoSFBatch``,
  ``calculatePathInfoEmu``, ``calculatePathInfoEmuBatch`` ...) plus the batched
  ``score_many(V, E, paths)`` the 64k-candidate configuration uses.  It relies on the host
  agent's own ``pathToTrajPoints`` (``GraceRIGV3.py:396-427``), which stays Python.

What the reference does per candidate -- k (sequential) or one (log-det) full GPy refits of
the (N+k)-point model -- becomes one shared L^-1 K* contraction plus k x k Schur-complement
work per candidate; targets of appended points are zero, so only covariances matter.
"""
import numpy as np

from . import _lib as L
from .core import GPCore, to_x4
from .gp_models import GPRegression, GPyMultiOutputWrapper

COND, PRE = 1, 2          # row-mask bits of gpc_ig_seq
LOG_DBL_MIN = np.log(np.nextafter(0, 1))   # below this np.linalg.det underflows to 0
LOG_DBL_MAX = np.log(np.finfo(float).max)  # above this it overflows to inf


def label_fidelity(var, fidLevs, bounded_top=False):
    """Fidelity index of trajectory points from their localisation variance
    (``GraceRIGV3.py:529-533``; ``bounded_top`` = the ``l3`` form of ``:508-512`` in which points
    above ``fidLevs[2]`` or exactly on a threshold get index 0 all the same)."""
    var = np.asarray(var, dtype=float).ravel()
    l1 = var < fidLevs[0]
    l2 = np.logical_and(var > fidLevs[0], var < fidLevs[1])
    return (l1 * 2 + l2 * 1).astype(float)


def _model_core(model):
    """Device core (factored on the model's current data/hypers) of a mirrored model."""
    if isinstance(model, GPyMultiOutputWrapper):
        model = model.gpy_model
    return model._ensure_factor(), model


def _rows(cands, mf):
    out = []
    for c in cands:
        c = np.asarray(c, dtype=float)
        if c.size == 0:
            out.append(np.zeros((0, 4)))
        elif mf:
            out.append(to_x4(c[:, :3], c[:, 3]))
        else:
            out.append(to_x4(c[:, :3]))
    return out


def seq_info_gain(model, cands, sig_n, pred_fid=-1, first_preadded=False, masks=None):
    """I_c = sum_i log(1 + sigma^2(x_i | data u x_<i) / sig_n) for every candidate.

    model: ``GPRegression`` (rows (k,3)) or the multi-fidelity wrapper (rows (k,4), fidelity last;
    every point is queried at ``pred_fid`` when >= 0, ``calculatePathInfoEmu`` uses 0).
    first_preadded: the first point is appended before it is predicted (``GraceRIGV3.py:454-455``).
    masks: optional list of per-row uint8 masks (bit0 conditioned-on, bit1 pre-appended).
    Returns (I (C,), argmax)."""
    core, m = _model_core(model)
    mf = not isinstance(m, GPRegression)
    rows, offs = GPCore._ragged(_rows(cands, mf))
    mask = None
    if masks is not None:
        mask = np.zeros(rows.shape[0], dtype=np.uint8)
        if offs[-1]:
            mask[:offs[-1]] = np.concatenate([np.asarray(x, np.uint8).ravel() for x in masks if len(x)])
    return core.ig_seq(rows, offs, sig_n, pred_fid if mf else -1,
                       L.IG_FIRST_PREADDED if first_preadded else 0, mask)


def logdet_info_gain(model, grid, cands, clip=False):
    """Raw I_c = 0.5 (logdet S_prior(grid) - logdet S_post(grid | data u X_c)) for every candidate,
    S = noise-inclusive predictive covariance (clip=True: clipped element-wise at 1e-10 like emukit's
    ``predict_covariance``).  Returns (I (C,), logdet_prior, argmax)."""
    core, m = _model_core(model)
    mf = not isinstance(m, GPRegression)
    grid = np.asarray(grid, dtype=float)
    g4 = to_x4(grid[:, :3], grid[:, 3]) if mf else to_x4(grid[:, :3])
    rows, offs = GPCore._ragged(_rows(cands, mf))
    return core.ig_logdet(g4, rows, offs, clip=clip)


def selfgrid_info_gain(model, cands, pred_fid=2, clip=True):
    """``calculatePathInfoEmu2``: I_c = 0.5 (logdet K(Xp) - logdet S_post(Xp | data u X_c)) with the
    grid Xp = the candidate's own points at ``pred_fid``.  Returns (I (C,), argmax)."""
    core, m = _model_core(model)
    rows, offs = GPCore._ragged(_rows(cands, True))
    return core.ig_selfgrid(rows, offs, pred_fid, clip)


def _guarded_sf_batch(I_raw, logdet_prior, G, prior_fallback):
    """The ``det == 0`` / ``isinf`` guards of ``calcPathInfoSFBatch``
    (``PhysicalExperimentCode/GraceRIGV3.py:583-596``) re-expressed on log-determinants:
    ``np.linalg.det`` returns 0 below exp(-744.4) and inf above exp(709.8)."""
    ldp = logdet_prior
    if ldp < LOG_DBL_MIN:
        ldp = prior_fallback
    elif ldp > LOG_DBL_MAX:
        ldp = np.inf
    ld_post = logdet_prior - 2.0 * np.asarray(I_raw)
    ld_post = np.where(ld_post < LOG_DBL_MIN, 0.0, np.where(ld_post > LOG_DBL_MAX, np.inf, ld_post))
    with np.errstate(invalid="ignore"):
        I = np.maximum(0.5 * (ldp - ld_post), 0.0)
    I = np.where(np.isinf(I), 0.0, I)
    return I


class InfoGainOperators:
    """Mixin for a ``GraceAgent``-like object.  Needs: ``pathToTrajPoints(V, E, path, dense=...,
    withVar=...)``, ``fidLevs``, ``fieldGrid``, ``sfgp`` and / or ``mfgp`` (mirrored models), and
    the cache slot ``logDetPrior`` the planner resets per ``plan()``
    (``PhysicalExperimentCode/GraceRIGV3.py:1314``)."""

    logDetPrior = None

    # -- path -> points (host, reference code) ---------------------------------------------
    def _sf_points(self, V, E, path, dense):
        return self.pathToTrajPoints(V, E, path, dense=dense)[:, :3]

    def _mf_points(self, V, E, path, dense, bounded_top=True):
        p = self.pathToTrajPoints(V, E, path, dense=dense, withVar=True)
        return np.hstack([p[:, :3], label_fidelity(p[:, -1], self.fidLevs, bounded_top)[:, None]])

    # -- sequential variants ------------------------------------------------------------------
    def calcPathInfoSF2_many(self, V, E, paths, dense=True):
        pts = [self._sf_points(V, E, p, dense) for p in paths]
        sig_n = float(self.sfgp.Gaussian_noise.variance[0])
        I, _ = seq_info_gain(self.sfgp, pts, sig_n, first_preadded=True)
        I = np.array(I, dtype=float)
        for c, p in enumerate(pts):  # reference: `if 0 in X.shape: return -np.inf` with X = pnts[1:]
            if p.shape[0] < 2:
                I[c] = -np.inf
        return I

    def calcPathInfoSF2(self, V, E, path, dense=True):
        """``GraceRIGV3.py:443-466``."""
        return float(self.calcPathInfoSF2_many(V, E, [path], dense)[0])

    calcPathInfoSF3 = calcPathInfoSF2  # ``Phys/GraceRIGV3.py:471-496``: same arithmetic, cached copy

    # A conditioning set far away from everything: zero covariance with every point (exp(-huge) == 0.0), so a
    # model holding only this row behaves exactly like an EMPTY training set.
    _FAR = 1.0e9

    def _window_model(self, slot, model, X, keep, set_data):
        """Cached copy of ``model`` conditioned on the rows of X inside the window (targets zero); a single
        far-away row stands in for an empty window."""
        inner = getattr(model, "gpy_model", model)
        stamp = (inner.param_array.tobytes(), inner._data_version)
        if getattr(self, "_%s_window_stamp" % slot, None) != stamp:
            wm = model.copy()
            Xw = X[keep]
            if Xw.shape[0] == 0:
                Xw = np.full((1, X.shape[1]), self._FAR)
                if X.shape[1] == 4:
                    Xw[0, 3] = 0.0
            getattr(wm, set_data)(Xw, np.zeros((Xw.shape[0], 1)))
            setattr(self, "_%s_window_model" % slot, wm)
            setattr(self, "_%s_window_stamp" % slot, stamp)
        return getattr(self, "_%s_window_model" % slot)

    def _sf_windowed_many(self, V, E, paths, dense, first_windowed):
        """Windowed sequential SF variants.  Row j of a path (j = 0 is the path's first point) is scored
        against data + points 0..j -- the point itself included -- as long as that set has <= 100 rows
        (N + 1 + j), and against its rows with x < 3 lx and y < 3 ly afterwards (the whole set when none
        qualifies); row 0 always sees the whole set.  ``first_windowed=False``: ``calcPathInfoSF``
        (``GraceRIGV3.py:468-503``).  ``first_windowed=True``: ``calcPathInfoSF4``
        (``Phys/GraceRIGV3.py:498-534``) -- row 0 is scored against the window whatever the size, without
        the empty-window fallback, and is appended twice when it lies inside it (``:512-513``)."""
        gp = self.sfgp
        pts = [self._sf_points(V, E, p, dense) for p in paths]
        sig_n = float(gp.Gaussian_noise.variance[0])
        lx, ly = [float(v) for v in gp.kern.lengthscale[:2]]
        inwin = lambda X: np.logical_and(X[:, 0] < 3 * lx, X[:, 1] < 3 * ly)
        N = gp.X.shape[0]
        out = np.full(len(pts), -np.inf)
        live = [c for c, p in enumerate(pts) if p.shape[0] >= 2]
        if not live:
            return out
        keep = inwin(gp.X)
        wm = self._window_model("sf", gp, gp.X, keep, "set_XY")
        P = [pts[c] for c in live]
        wmask = [np.where(inwin(p), COND | PRE, 0).astype(np.uint8) for p in P]
        s0 = max(1, 100 - N)                        # first row whose set (N + 1 + j rows) exceeds 100
        if keep.any():
            sw = [min(s0, len(p)) for p in P]
        else:                                       # empty window until the first path point falls inside it
            first_in = [int(np.argmax(m > 0)) if (m > 0).any() else len(m) for m in wmask]
            sw = [min(max(s0, f), len(p)) for f, p in zip(first_in, P)]
        allpre = lambda n: np.full(n, COND | PRE, dtype=np.uint8)
        I_full, _ = seq_info_gain(gp, [p[:s] for p, s in zip(P, sw)], sig_n, masks=[allpre(s) for s in sw])
        I_all, _ = seq_info_gain(wm, P, sig_n, masks=wmask)
        I_head, _ = seq_info_gain(wm, [p[:s] for p, s in zip(P, sw)], sig_n, masks=[m[:s] for m, s in zip(wmask, sw)])
        total = np.asarray(I_full) + np.asarray(I_all) - np.asarray(I_head)
        if first_windowed:
            heads = [p[:1] for p in P]
            dup = [np.concatenate((h, h)) for h in heads]            # window copy (if inside) + the explicit append
            m2 = [np.array([m[0] & COND, COND | PRE], dtype=np.uint8) for m in wmask]
            I2, _ = seq_info_gain(wm, dup, sig_n, masks=m2)
            I1, _ = seq_info_gain(wm, heads, sig_n, masks=[np.array([m[0] & COND], dtype=np.uint8) for m in wmask])
            I0, _ = seq_info_gain(gp, heads, sig_n, masks=[allpre(1) for _ in heads])
            total += np.asarray(I2) - np.asarray(I1) - np.asarray(I0)
        out[live] = total
        return out

    def calcPathInfoSF_many(self, V, E, paths, dense=True):
        return self._sf_windowed_many(V, E, paths, dense, first_windowed=False)

    def calcPathInfoSF(self, V, E, path, dense=True):
        return float(self.calcPathInfoSF_many(V, E, [path], dense)[0])

    def calcPathInfoSF4_many(self, V, E, paths, dense=True):
        return self._sf_windowed_many(V, E, paths, dense, first_windowed=True)

    def calcPathInfoSF4(self, V, E, path, dense=True):
        return float(self.calcPathInfoSF4_many(V, E, [path], dense)[0])

    def calculatePathInfoEmu_many(self, V, E, paths, dense=False, sig_index=-1, windowed=True):
        """``GraceRIGV3.py:525-562`` (``sig_index=-3``) / ``Phys/GraceRIGV3.py:641-678`` (``-1``).
        Point j is predicted (at fidelity 0) from data + points 0..j-1 while data + points 0..j have <= 100
        rows; afterwards the reference keeps only rows with x < 5 lx and y < 5 ly of data + points 0..j --
        the point being predicted is then part of its own conditioning set (``tempX`` is cut from ``allX``
        after the append, ``:549-553``)."""
        pts = [self._mf_points(V, E, p, dense, bounded_top=False) for p in paths]
        gm = self.mfgp.gpy_model
        sig_n = float(gm.param_array[sig_index])
        N = gm.X.shape[0]
        s0 = max(0, 100 - N) if windowed else 1 << 30   # first windowed point
        sw = [min(s0, len(p)) for p in pts]
        I_full, _ = seq_info_gain(self.mfgp, [p[:s] for p, s in zip(pts, sw)], sig_n, pred_fid=0)
        if all(s == len(p) for p, s in zip(pts, sw)):
            return np.asarray(I_full)
        lx, ly = [float(v) for v in gm.kern.kernels[0].lengthscale[:2]]
        inwin = lambda X: np.logical_and(X[:, 0] < 5 * lx, X[:, 1] < 5 * ly)
        keep = inwin(gm.X)
        masks = [np.where(inwin(p), COND | PRE, 0).astype(np.uint8) for p in pts]
        if not keep.any():
            for m, s in zip(masks, sw):
                if s < len(m) and not (m[:s + 1] > 0).any():
                    raise ValueError("empty conditioning window (the reference's set_data fails on it as well)")
        wmodel = self._window_model("mf", self.mfgp, gm.X, keep, "set_data")
        I_all, _ = seq_info_gain(wmodel, pts, sig_n, pred_fid=0, masks=masks)
        I_head, _ = seq_info_gain(wmodel, [p[:s] for p, s in zip(pts, sw)], sig_n, pred_fid=0,
                                  masks=[m[:s] for m, s in zip(masks, sw)])
        return np.asarray(I_full) + np.asarray(I_all) - np.asarray(I_head)

    def calculatePathInfoEmu(self, V, E, path, dense=False, sig_index=-1):
        return float(self.calculatePathInfoEmu_many(V, E, [path], dense, sig_index)[0])

    # -- log-det variants ---------------------------------------------------------------------
    def calcPathInfoSFBatch_many(self, V, E, paths, dense=True):
        """``Phys/GraceRIGV3.py:571-597`` for many paths, each scored against the agent's current
        data (the cumulative append of the reference's cached copy, ``:590``, is reproduced by the
        single-path ``calcPathInfoSFBatch`` only)."""
        pts = [self._sf_points(V, E, p, dense)[1:] for p in paths]
        I_raw, ldp, _ = logdet_info_gain(self.sfgp, self.fieldGrid, pts)
        G = np.asarray(self.fieldGrid).shape[0]
        fallback = G * np.log(float(self.sfgp.kern.variance[0]) + float(self.sfgp.Gaussian_noise.variance[0]))
        if self.logDetPrior is None:
            self.logDetPrior = ldp if LOG_DBL_MIN <= ldp <= LOG_DBL_MAX else (fallback if ldp < LOG_DBL_MIN else np.inf)
        return _guarded_sf_batch(I_raw, ldp, G, fallback)

    # True: the single-path ``calcPathInfoSFBatch`` / ``calculatePathInfoEmuBatch`` reproduce the reference's cached
    # model copies ``sfgp2`` / ``mfgp2``: made once, they keep the HYPER-PARAMETERS of that moment, and the SF copy's
    # data GROW across the calls of one ``plan()`` (see there).  The batched ``*_many`` / ``score_many`` calls always
    # score every path against the agent's current model and data only.
    reference_quirks = True

    def calcPathInfoSFBatch(self, V, E, path, dense=True):
        """``Phys/GraceRIGV3.py:571-597``.  The reference scores on a cached copy ``sfgp2`` that is reset to the
        agent's data only while ``logDetPrior`` is None (once per ``plan()``, ``:581-582,1314``) and appends the
        path's points to the COPY's data on every call (``:590``), so call c of a plan is conditioned on the
        points of calls 0..c-1 as well.  With ``reference_quirks`` (default) this is reproduced: one device
        refit + grid covariance + log-det per call; otherwise the path is scored independently."""
        if not self.reference_quirks:
            return float(self.calcPathInfoSFBatch_many(V, E, [path], dense)[0])
        X = self._sf_points(V, E, path, dense)[1:]
        grid = np.asarray(self.fieldGrid, dtype=float)
        G = grid.shape[0]
        m = getattr(self, "_sfb_model", None)
        if m is None:
            m = self._sfb_model = self.sfgp.copy()

        def grid_logdet():
            _, K = m.predict(grid, full_cov=True)
            return m._ensure_factor().spd_stats(K)[2]

        if self.logDetPrior is None:
            m.set_XY(self.sfgp.X, self.sfgp.Y)
            ld = grid_logdet()
            if ld < LOG_DBL_MIN:        # det == 0 -> det of (variance + noise) I
                ld = G * np.log(float(m.kern.variance[0]) + float(m.Gaussian_noise.variance[0]))
            elif ld > LOG_DBL_MAX:
                ld = np.inf
            self.logDetPrior = ld
        m.set_XY(np.concatenate((m.X, X)), np.concatenate((m.Y, np.zeros((X.shape[0], 1)))))
        ld = grid_logdet()
        ld = 0.0 if ld < LOG_DBL_MIN else (np.inf if ld > LOG_DBL_MAX else ld)
        with np.errstate(invalid="ignore"):
            I = max(0.5 * (self.logDetPrior - ld), 0)
        return 0.0 if np.isinf(I) else float(I)

    def calculatePathInfoEmuBatch_many(self, V, E, paths, dense=False):
        """``Phys/GraceRIGV3.py:599-618``: grid at fidelity 2, no guards, no clamp; both covariances come
        from emukit's ``predict_covariance``, i.e. clipped element-wise at 1e-10 (reproduced on the device)."""
        pts = [self._mf_points(V, E, p, dense, bounded_top=True) for p in paths]
        grid = np.asarray(self.fieldGrid, dtype=float)
        grid4 = np.hstack([grid[:, :3], 2 * np.ones((grid.shape[0], 1))])
        I_raw, ldp, _ = logdet_info_gain(self.mfgp, grid4, pts, clip=True)
        if self.logDetPrior is None:
            self.logDetPrior = ldp
        return np.asarray(I_raw)

    def calculatePathInfoEmuBatch(self, V, E, path, dense=False):
        """``Phys/GraceRIGV3.py:599-618``.  The reference scores on a cached copy ``mfgp2`` made the first time the
        operator runs (``:608-609``): its DATA are reset to the agent's on every call (``:611``) but its
        HYPER-PARAMETERS stay the ones the agent's model had when the copy was made, and ``logDetPrior`` is the one of
        the first call of the ``plan()``.  With ``reference_quirks`` (default) both are reproduced; otherwise (and in
        the batched ``*_many`` / ``score_many`` calls) the agent's current model is used."""
        if not self.reference_quirks:
            return float(self.calculatePathInfoEmuBatch_many(V, E, [path], dense)[0])
        m = getattr(self, "_mfb_model", None)
        if m is None:
            m = self._mfb_model = self.mfgp.copy()
        m.set_data(self.mfgp.X, self.mfgp.Y)
        pts = self._mf_points(V, E, path, dense, bounded_top=True)
        grid = np.asarray(self.fieldGrid, dtype=float)
        grid4 = np.hstack([grid[:, :3], 2 * np.ones((grid.shape[0], 1))])
        I_raw, ldp, _ = logdet_info_gain(m, grid4, [pts], clip=True)
        if self.logDetPrior is None:
            self.logDetPrior = ldp
        # I = 1/2 (logDetPrior - logdet posterior), the posterior log-det recovered from the raw gain
        return float(I_raw[0] + 0.5 * (self.logDetPrior - ldp))

    def calculatePathInfoEmu2_many(self, V, E, paths, dense=False):
        """``GraceRIGV3.py:505-523``: grid = the candidate's own points at fidelity 2, prior = the
        kernel matrix, posterior = emukit ``predict_covariance`` (clipped at 1e-10)."""
        pts = [self._mf_points(V, E, p, dense, bounded_top=True) for p in paths]
        I, _ = selfgrid_info_gain(self.mfgp, pts, pred_fid=2, clip=True)
        return np.asarray(I)

    def calculatePathInfoEmu2(self, V, E, path, dense=False):
        return float(self.calculatePathInfoEmu2_many(V, E, [path], dense)[0])

    # -- batched operator slot ------------------------------------------------------------------
    def score_many(self, V, E, paths, operator="calculatePathInfoEmuBatch", **kw):
        """Score a list of candidate paths in one device pass; returns (I (C,), argmax)."""
        I = getattr(self, operator + "_many")(V, E, paths, **kw)
        finite = np.where(np.isnan(I), -np.inf, I)
        return I, (int(np.argmax(finite)) if len(I) else -1)
