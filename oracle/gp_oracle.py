"""NumPy/SciPy restatement of the reference's dense GP inference path.

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  Nothing under
``gpcore`` (the product) may import this module.

What is restated, and from where (paths relative to the reference checkout):

* SE-ARD kernel                      ``NIGP.py:11-20``  (GPy ``RBF.K``; Gram-trick distances)
* posterior mean + input gradients   ``NIGP.py:29-65``
* NIGP negative log marginal lik.    ``NIGP.py:130-165`` (+ ``safe_obj`` ``:119-123``)
* NIGP.predict                       ``NIGP.py:269-333``
* SF-GP (GPy ``GPRegression``)       call sites ``GPTrainers.py:80-98,116-117``,
                                     ``PhysicalExperimentCode/GraceRIGV3.py:446-597``
* MF-GP (emukit linear AR1 model)    call sites ``GPTrainers.py:59-74,119-120``,
                                     ``GraceRIGV3.py:505-562``
* information-gain operators         ``GraceRIGV3.py:443-562``,
                                     ``PhysicalExperimentCode/GraceRIGV3.py:571-678``,
                                     ``informationGainTest.py:1-52``

GPy / emukit are third-party, un-vendored and un-pinned (PARITY UNPINNED at that
boundary); each quirk recalled from their published source is an explicit keyword so
it can be tested on its own:

    gram=True        distances by the Gram identity, clipped at 0 (GPy ``Stationary``)
    jitter=1e-8      ``exact_gaussian_inference`` adds noise + 1e-8 to the diagonal
    include_noise    ``predict`` adds the likelihood variance (GPy default)
    clip_diag=1e-15  clip of the latent marginal variance
    clip_cov=1e-10   emukit ``predict_covariance`` element-wise clip
"""
import numpy as np
from scipy.linalg import cho_factor, cho_solve, solve_triangular

SQRT3 = np.sqrt(3.0)
GPY_JITTER = 1e-8
KIND_RBF = 0
KIND_MAT32 = 1


# --------------------------------------------------------------------------------------
# kernels
# --------------------------------------------------------------------------------------
def scaled_sqdist(X1, X2, ls, gram=True, same=False):
    """r^2 = sum_d ((x_d - x'_d)/l_d)^2.  gram=True follows GPy ``_unscaled_dist``."""
    A = np.asarray(X1, float) / ls
    B = np.asarray(X2, float) / ls
    if gram:
        r2 = -2.0 * A.dot(B.T) + (np.sum(A * A, 1)[:, None] + np.sum(B * B, 1)[None, :])
        if same:
            r2[np.diag_indices(A.shape[0])] = 0.0
        return np.clip(r2, 0.0, np.inf)
    # direct differences, one input dimension at a time: (n1, n2) temporaries instead of (n1, n2, D), so that
    # the BASELINE sizes (N = 16384) fit in host memory; same summation order as a sum over the last axis
    r2 = np.zeros((A.shape[0], B.shape[0]))
    for d in range(A.shape[1]):
        diff = A[:, d][:, None] - B[:, d][None, :]
        diff *= diff
        r2 += diff
    return r2


def k_stationary(X1, X2, variance, ls, kind=KIND_RBF, gram=True, same=False):
    r2 = scaled_sqdist(X1, X2, np.asarray(ls, float), gram=gram, same=same)
    if kind == KIND_RBF:
        return variance * np.exp(-0.5 * r2)
    r = np.sqrt(r2)
    return variance * (1.0 + SQRT3 * r) * np.exp(-SQRT3 * r)


def SE_ARD_kernel(X1, X2, lengthscales, sigma_f, gram=True):
    """``NIGP.py:11-20``: sigma_f is passed to GPy as the kernel *variance*."""
    ls = np.sqrt(1.0 / (1.0 / np.asarray(lengthscales, float) ** 2))  # inv_l=True round trip
    return k_stationary(X1, X2, sigma_f, ls, KIND_RBF, gram=gram)


def ar1_coeff(rho, i, m):
    """prod_{l=m}^{i-1} rho_l (emukit ``LinearMultiFidelityKernel``)."""
    return float(np.prod(rho[m:i])) if i > m else 1.0


def k_ar1(X4a, X4b, variances, ls, rho, kind=KIND_RBF, gram=True, same=False):
    """Kennedy-O'Hagan AR1 covariance between fidelity-indexed rows (x,y,z,fid).

    K[(x,i),(x',j)] = sum_{m<=min(i,j)} (prod_{l=m}^{i-1} rho_l)(prod_{l=m}^{j-1} rho_l) k_m(x,x')
    """
    X4a = np.asarray(X4a, float)
    X4b = np.asarray(X4b, float)
    fa = X4a[:, -1].astype(int)
    fb = X4b[:, -1].astype(int)
    F = len(variances)
    rho = np.asarray(rho, float)
    base = [k_stationary(X4a[:, :-1], X4b[:, :-1], variances[m], ls[m], kind, gram=gram, same=same)
            for m in range(F)]
    K = np.zeros((X4a.shape[0], X4b.shape[0]))
    for i in range(F):
        for j in range(F):
            mask = (fa == i)[:, None] & (fb == j)[None, :]
            if not mask.any():
                continue
            blk = np.zeros_like(K)
            for m in range(min(i, j) + 1):
                blk += ar1_coeff(rho, i, m) * ar1_coeff(rho, j, m) * base[m]
            K[mask] = blk[mask]
    return K


def k_ar1_diag(X4, variances, rho):
    f = np.asarray(X4)[:, -1].astype(int)
    rho = np.asarray(rho, float)
    out = np.zeros(len(f))
    for i in range(len(variances)):
        v = sum(ar1_coeff(rho, i, m) ** 2 * variances[m] for m in range(i + 1))
        out[f == i] = v
    return out


# --------------------------------------------------------------------------------------
# hyper-parameter vector layouts (the reference's ``param_array`` orders)
# --------------------------------------------------------------------------------------
def split_sf_params(p):
    """[variance, lx, ly, lz, noise_var]  (header at PhysicalExperimentCode/...SFGP.py:620)."""
    p = np.asarray(p, float)
    return p[0], p[1:-1], p[-1]


def split_mf_params(p, F, D=3):
    """[var_0,l_0(D), ..., var_{F-1},l_{F-1}(D), rho(F-1), noise (1 | F)]
    (header at PhysicalExperimentCode/GraceExplorationExperiments_MFGP.py:670)."""
    p = np.asarray(p, float)
    per = D + 1
    variances = np.array([p[m * per] for m in range(F)])
    ls = np.array([p[m * per + 1:(m + 1) * per] for m in range(F)])
    rho = p[F * per:F * per + F - 1]
    noise = p[F * per + F - 1:]
    if noise.size == 1:
        noise = np.repeat(noise, F)
    assert noise.size == F
    return variances, ls, rho, noise


# --------------------------------------------------------------------------------------
# exact Gaussian inference
# --------------------------------------------------------------------------------------
class Factor:
    """Cholesky state of Ky = K + diag(noise) (+ jitter)."""

    def __init__(self, Ky, y):
        self.L = np.linalg.cholesky(Ky)  # raises LinAlgError when not PD (reference convention)
        y = np.asarray(y, float).reshape(-1)
        self.alpha = cho_solve((self.L, True), y)
        self.logdet = 2.0 * np.sum(np.log(np.diag(self.L)))
        self.N = Ky.shape[0]
        self.nlml = 0.5 * float(y @ self.alpha) + 0.5 * self.logdet + 0.5 * self.N * np.log(2 * np.pi)

    def half_solve(self, Kx):
        """L^{-1} Kx  (Kx is N x M)."""
        return solve_triangular(self.L, Kx, lower=True, check_finite=False)


class SFGP:
    """GPy ``GPRegression``-equivalent arithmetic (zero mean, Gaussian likelihood)."""

    def __init__(self, X, Y, params, kind=KIND_RBF, gram=True, jitter=GPY_JITTER):
        self.kind, self.gram, self.jitter = kind, gram, jitter
        self.params = np.asarray(params, float).copy()
        self.set_XY(X, Y)

    def kern(self, A, B, same=False):
        v, ls, _ = split_sf_params(self.params)
        return k_stationary(A, B, v, ls, self.kind, gram=self.gram, same=same)

    def set_XY(self, X, Y):
        self.X = np.asarray(X, float)
        self.Y = np.asarray(Y, float).reshape(-1, 1)
        _, _, noise = split_sf_params(self.params)
        Ky = self.kern(self.X, self.X, same=True)
        Ky[np.diag_indices_from(Ky)] += noise + self.jitter
        self.f = Factor(Ky, self.Y)

    def predict(self, Xs, full_cov=False, include_noise=True, clip_diag=1e-15):
        v, _, noise = split_sf_params(self.params)
        Xs = np.asarray(Xs, float)
        Kx = self.kern(self.X, Xs)
        mu = Kx.T @ self.f.alpha
        tmp = self.f.half_solve(Kx)
        if full_cov:
            var = self.kern(Xs, Xs, same=True) - tmp.T @ tmp
            if include_noise:
                var = var + noise * np.eye(Xs.shape[0])
            return mu[:, None], var
        var = v - np.sum(tmp * tmp, 0)
        if clip_diag is not None:
            var = np.clip(var, clip_diag, np.inf)
        if include_noise:
            var = var + noise
        return mu[:, None], var[:, None]


class MFGP:
    """emukit ``GPyLinearMultiFidelityModel`` + ``GPyMultiOutputWrapper`` arithmetic."""

    def __init__(self, X4, Y, params, F=3, kind=KIND_RBF, gram=True, jitter=GPY_JITTER):
        self.F, self.kind, self.gram, self.jitter = F, kind, gram, jitter
        self.params = np.asarray(params, float).copy()
        self.set_data(X4, Y)

    def kern(self, A, B, same=False):
        variances, ls, rho, _ = split_mf_params(self.params, self.F)
        return k_ar1(A, B, variances, ls, rho, self.kind, gram=self.gram, same=same)

    def noise_of(self, X4):
        _, _, _, noise = split_mf_params(self.params, self.F)
        return noise[np.asarray(X4)[:, -1].astype(int)]

    def set_data(self, X4, Y):
        self.X = np.asarray(X4, float)
        self.Y = np.asarray(Y, float).reshape(-1, 1)
        Ky = self.kern(self.X, self.X, same=True)
        Ky[np.diag_indices_from(Ky)] += self.noise_of(self.X) + self.jitter
        self.f = Factor(Ky, self.Y)

    def predict(self, X4s, include_noise=True, clip_diag=1e-15):
        variances, _, rho, _ = split_mf_params(self.params, self.F)
        X4s = np.asarray(X4s, float)
        Kx = self.kern(self.X, X4s)
        mu = Kx.T @ self.f.alpha
        tmp = self.f.half_solve(Kx)
        var = k_ar1_diag(X4s, variances, rho) - np.sum(tmp * tmp, 0)
        if clip_diag is not None:
            var = np.clip(var, clip_diag, np.inf)
        if include_noise:
            var = var + self.noise_of(X4s)
        return mu[:, None], var[:, None]

    def predict_covariance(self, X4s, include_noise=True, clip_cov=1e-10):
        X4s = np.asarray(X4s, float)
        Kx = self.kern(self.X, X4s)
        tmp = self.f.half_solve(Kx)
        cov = self.kern(X4s, X4s, same=True) - tmp.T @ tmp
        if include_noise:
            cov = cov + np.diag(self.noise_of(X4s))
        if clip_cov is not None:
            cov = np.clip(cov, clip_cov, np.inf)
        return cov


class MFGPCached(MFGP):
    """``MFGP`` whose ``set_data`` re-uses the covariance block of the FIRST training set it saw when the new set
    extends it (the literal refit loops of the information-gain operators append a few rows per refit).  Only the
    ASSEMBLY is incremental -- kernel entries do not depend on one another, so the matrix is identical -- every refit
    still factors the whole extended matrix from scratch, as the reference does.  Makes the loops affordable at
    N = 4096 (0.3 s instead of 5 s per refit)."""

    def set_data(self, X4, Y):
        X4 = np.asarray(X4, float)
        base = getattr(self, "_X0", None)
        if base is None:
            self._X0 = X4.copy()
            self._K0 = self.kern(X4, X4, same=True)
            base = self._X0
        n0 = base.shape[0]
        if X4.shape[0] >= n0 and np.array_equal(X4[:n0], base):
            n = X4.shape[0]
            Ky = np.empty((n, n))
            Ky[:n0, :n0] = self._K0
            if n > n0:
                Kn = self.kern(X4[n0:], X4)
                Ky[n0:, :] = Kn
                Ky[:n0, n0:] = Kn[:, :n0].T
        else:
            Ky = self.kern(X4, X4, same=True)
        self.X = X4
        self.Y = np.asarray(Y, float).reshape(-1, 1)
        Ky[np.diag_indices_from(Ky)] += self.noise_of(self.X) + self.jitter
        self.f = Factor(Ky, self.Y)


# --------------------------------------------------------------------------------------
# NIGP restatement (the pinned reference is NIGP.py itself, imported through gpy_shim)
# --------------------------------------------------------------------------------------
def compute_post_mean_and_gradients(X, y, ls, sigma_f, sigma_y, noise_diag=None, gram=True):
    """``NIGP.py:29-65`` without the Python loop over i."""
    X = np.asarray(X, float)
    y = np.asarray(y, float).reshape(-1)
    N, D = X.shape
    nd = np.zeros(N) if noise_diag is None else np.asarray(noise_diag, float)
    K = SE_ARD_kernel(X, X, ls, sigma_f, gram=gram)
    f = Factor(K + np.diag(sigma_y ** 2 + nd), y)
    inv_ls2 = 1.0 / np.asarray(ls, float) ** 2
    W = K * f.alpha[None, :]                       # W_ij = K_ij alpha_j
    grads = -(X * W.sum(1)[:, None] - W @ X) * inv_ls2[None, :]
    return K @ f.alpha, grads


def nigp_nlml(log_hyp, X, y, grad_fixed, extra=None, gram=True):
    """``NIGP.py:130-165`` (returns 1e25 on a non-PD matrix, like the reference)."""
    X = np.asarray(X, float)
    y = np.asarray(y, float).reshape(-1)
    N, D = X.shape
    ls = np.exp(log_hyp[:D]); sf = np.exp(log_hyp[D]); sy = np.exp(log_hyp[D + 1]); sx = np.exp(log_hyp[D + 2:])
    v = np.sum(grad_fixed ** 2 * sx[None, :] ** 2, axis=1)
    if extra is not None:
        v = v + extra
    K = SE_ARD_kernel(X, X, ls, sf, gram=gram)
    try:
        f = Factor(K + np.diag(sy ** 2 + v) + np.eye(N) * 1e-8, y)
    except np.linalg.LinAlgError:
        return 1e25
    return float(f.nlml)


def nigp_predict(X, y, ls, sigma_f, sigma_y, noise_diag, Xs, Xs_input_noise=None,
                 return_var=True, return_cov=False, gram=True):
    """``NIGP.py:269-333`` -- diagonal computed without the M x M matrix when return_cov=False."""
    X = np.asarray(X, float); y = np.asarray(y, float).reshape(-1); Xs = np.asarray(Xs, float)
    ls = np.asarray(ls, float)
    K = SE_ARD_kernel(X, X, ls, sigma_f, gram=gram)
    obs = sigma_y ** 2 + (noise_diag if noise_diag is not None else 0.0)
    f = Factor(K + np.diag(obs * np.ones(X.shape[0])), y)
    Kxs = SE_ARD_kernel(Xs, X, ls, sigma_f, gram=gram)
    mean = Kxs @ f.alpha
    if not (return_var or return_cov):
        return mean
    tmp = f.half_solve(Kxs.T)
    if return_cov:
        cov = SE_ARD_kernel(Xs, Xs, ls, sigma_f, gram=gram) - tmp.T @ tmp
        diag = None
    else:
        diag = sigma_f - np.sum(tmp * tmp, 0)
    if Xs_input_noise is not None:
        M, D = Xs.shape
        W = Kxs * f.alpha[None, :]
        grads = -(Xs * W.sum(1)[:, None] - W @ X) / ls[None, :] ** 2
        sx = np.asarray(Xs_input_noise, float)
        if sx.ndim == 1 and sx.size == D:
            sx = sx[None, :]
        elif sx.shape != grads.shape:
            raise ValueError("Xs_input_noise must be scalar, shape (D,) or (M,D)")
        vstar = np.sum(grads ** 2 * sx ** 2, axis=1)
        if return_cov:
            cov = cov + np.diag(vstar)
        else:
            diag = diag + vstar
    if return_cov:
        return mean, cov + np.eye(cov.shape[0]) * 1e-12
    return mean, np.maximum(diag + 1e-12, 1e-12)


def nigp_fit(X, y, n_restarts=3, iters=3, maxiter_opt=200, gram=True):
    """``NIGP.py:191-260`` restated: median-distance initialisation (``:198-212``), alternation of (A) posterior-mean
    input gradients with the input-noise term off (``:218-225``) and (B) ``n_restarts`` L-BFGS-B runs (numerical
    gradients, bounds [1e-6, 1e6]) from ``log_hyp + 0.1 randn`` drawn from NumPy's GLOBAL generator (``:231-239``).
    Returns (get_params()-ordered hypers, log_hyp, noise_diag_train, best objective of the last round)."""
    from scipy.optimize import minimize
    X = np.asarray(X, float)
    y = np.asarray(y, float).flatten()
    N, D = X.shape
    pairwise = np.sqrt(np.maximum(0, np.sum((X[:, None, :] - X[None, :, :]) ** 2, axis=2)))
    med = np.median(pairwise[pairwise > 0]) if np.any(pairwise > 0) else 1.0
    sf0 = np.std(y) if np.std(y) > 0 else 1.0
    sx0 = np.ones(D) * 0.01 * np.std(X, axis=0)
    log_hyp = np.concatenate([np.log(np.ones(D) * (med if med > 0 else 1.0)), [np.log(sf0), np.log(0.1 * sf0)],
                              np.log(np.maximum(sx0, 1e-8))])
    zeros = np.zeros(N)
    grads, best_val = np.zeros((N, D)), 1e99

    def obj(lh, g):
        v = nigp_nlml(lh, X, y, g, zeros, gram=gram)
        return v if np.isfinite(v) else 1e20

    for _ in range(iters):
        _, grads = compute_post_mean_and_gradients(X, y, np.exp(log_hyp[:D]), np.exp(log_hyp[D]), np.exp(log_hyp[D + 1]),
                                                   gram=gram)
        best, best_val, res = None, 1e99, None
        for _r in range(n_restarts):
            init = log_hyp + 0.1 * np.random.randn(*log_hyp.shape)
            res = minimize(lambda lh: obj(lh, grads), init, method="L-BFGS-B",
                           bounds=[(np.log(1e-6), np.log(1e6))] * (2 * D + 2), options={"maxiter": maxiter_opt})
            if res.fun < best_val:
                best_val, best = res.fun, res
        log_hyp = res.x if best is None else best.x
    h = np.exp(log_hyp)
    sx = h[D + 2:]
    return np.hstack((sx, h[D], h[D + 1], h[:D])), log_hyp, np.sum(grads ** 2 * sx[None, :] ** 2, axis=1), best_val


# --------------------------------------------------------------------------------------
# information gain -- literal refit loops (ground truth) and Schur forms (what the GPU does)
# --------------------------------------------------------------------------------------
def ig_seq_sf_refit(gp, Xc, first_preadded=True):
    """``GraceRIGV3.py:443-466`` (calcPathInfoSF2) literally: one refit per point.

    The first point is appended to the training set *before* it is predicted
    (``:454-455``); the others are predicted, then appended (``:458-463``).
    """
    _, _, sig_n = split_sf_params(gp.params)
    X0, Y0 = gp.X.copy(), gp.Y.copy()
    Xc = np.asarray(Xc, float)
    I = 0.0
    try:
        for i in range(Xc.shape[0]):
            x = Xc[i:i + 1]
            if i == 0 and first_preadded:
                gp.set_XY(np.concatenate((gp.X, x)), np.concatenate((gp.Y, [[0.0]])))
                _, s = gp.predict(x)
                I += np.log(1 + s[0, 0] / sig_n)
                continue
            _, s = gp.predict(x)
            I += np.log(1 + s[0, 0] / sig_n)
            gp.set_XY(np.concatenate((gp.X, x)), np.concatenate((gp.Y, [[0.0]])))
    finally:
        gp.set_XY(X0, Y0)
    return I


def ig_seq_mf_refit(gp, X4c, sig_n, pred_fid=0):
    """Un-windowed core of ``GraceRIGV3.py:525-562`` (calculatePathInfoEmu): each point is
    predicted at fidelity ``pred_fid`` given data + earlier candidates, then appended with
    its own fidelity label; targets are zeros."""
    X0, Y0 = gp.X.copy(), gp.Y.copy()
    X4c = np.asarray(X4c, float)
    I = 0.0
    try:
        gp.set_data(X0, np.zeros_like(Y0))
        for i in range(X4c.shape[0]):
            x = X4c[i:i + 1]
            xp = x.copy(); xp[0, -1] = pred_fid
            _, s = gp.predict(xp)
            I += np.log(1 + s[0, 0] / sig_n)
            gp.set_data(np.concatenate((gp.X, x)), np.concatenate((gp.Y, [[0.0]])))
    finally:
        gp.set_data(X0, Y0)
    return I


def _cond_blocks(gp, Xa, Xb=None):
    """Latent posterior (cross-)covariance given the model's data: K_ab - Va^T Vb."""
    Va = gp.f.half_solve(gp.kern(gp.X, Xa))
    if Xb is None:
        return gp.kern(Xa, Xa, same=True) - Va.T @ Va
    Vb = gp.f.half_solve(gp.kern(gp.X, Xb))
    return gp.kern(Xa, Xb) - Va.T @ Vb


def ig_seq_schur(gp, Xc, noise_train, noise_pred, sig_n, Xpred=None, first_preadded=False,
                 jitter=GPY_JITTER):
    """Sequential IG from one Cholesky of the k x k Schur complement.

    noise_train[i]: likelihood variance of candidate i once appended (jitter added here);
    noise_pred[i]:  likelihood variance added by ``predict`` at the query row;
    Xpred: query rows if they differ from the appended rows (MF: fidelity ``pred_fid``).
    """
    Xc = np.asarray(Xc, float)
    k = Xc.shape[0]
    S = _cond_blocks(gp, Xc) + np.diag(np.asarray(noise_train, float) * np.ones(k) + jitter)
    if Xpred is None:
        Cq = S - np.diag(np.asarray(noise_train, float) * np.ones(k) + jitter)   # latent
        q = np.diag(Cq).copy()
        cross = Cq
    else:
        Xpred = np.asarray(Xpred, float)
        q = np.diag(_cond_blocks(gp, Xpred)).copy()
        cross = _cond_blocks(gp, Xc, Xpred)        # cross[j, i] = cov(f(c_j), f(p_i) | D)
    Ls = np.linalg.cholesky(S)
    I = 0.0
    npred = np.asarray(noise_pred, float) * np.ones(k)
    for i in range(k):
        upto = i + 1 if (i == 0 and first_preadded) else i
        if upto:
            w = solve_triangular(Ls[:upto, :upto], cross[:upto, i], lower=True, check_finite=False)
            lat = q[i] - w @ w
        else:
            lat = q[i]
        lat = max(lat, 1e-15)                      # GPy latent-variance clip
        I += np.log(1 + (lat + npred[i]) / sig_n)
    return I


def ig_logdet_refit(gp, grid, Xc, noise_c=None, clip_cov=None):
    """``PhysicalExperimentCode/GraceRIGV3.py:571-618`` literally (independent per call):
    0.5 (logdet Sigma_prior(grid | data) - logdet Sigma_post(grid | data u Xc)), both
    noise-inclusive predictive covariances, determinants via slogdet."""
    is_mf = isinstance(gp, MFGP)
    X0, Y0 = gp.X.copy(), gp.Y.copy()
    cov = (lambda g: gp.predict_covariance(g, clip_cov=clip_cov)) if is_mf else \
          (lambda g: gp.predict(g, full_cov=True)[1])
    setd = gp.set_data if is_mf else gp.set_XY
    try:
        prior = np.linalg.slogdet(cov(grid))[1]
        setd(np.concatenate((X0, Xc)), np.concatenate((Y0, np.zeros((len(Xc), 1)))))
        post = np.linalg.slogdet(cov(grid))[1]
    finally:
        setd(X0, Y0)
    return 0.5 * (prior - post)


def ig_logdet_schur(gp, grid, Xc, noise_grid, noise_train, jitter=GPY_JITTER):
    """Same quantity by the matrix-determinant lemma (k x k work per candidate):
    logdet(Sg - B S^{-1} B^T) = logdet Sg + logdet(S - B^T Sg^{-1} B) - logdet S."""
    G = len(grid); k = len(Xc)
    Sg = _cond_blocks(gp, grid) + np.diag(np.asarray(noise_grid, float) * np.ones(G))
    S = _cond_blocks(gp, Xc) + np.diag(np.asarray(noise_train, float) * np.ones(k) + jitter)
    B = _cond_blocks(gp, grid, Xc)                       # G x k
    T = S - B.T @ cho_solve(cho_factor(Sg, lower=True), B)
    return 0.5 * (np.linalg.slogdet(S)[1] - np.linalg.slogdet(T)[1])


def ig_selfgrid_refit(gp, X4c, pred_fid=2, clip_cov=1e-10):
    """``GraceRIGV3.py:505-523`` (calculatePathInfoEmu2) literally: grid = the candidate's own
    points at fidelity ``pred_fid``; prior = kern.K(Xpred) (no noise, no data); posterior =
    predict_covariance after appending the candidate with zero targets; log(det) via slogdet."""
    X4c = np.asarray(X4c, float)
    Xp = X4c.copy(); Xp[:, -1] = pred_fid
    X0, Y0 = gp.X.copy(), gp.Y.copy()
    try:
        Kprior = gp.kern(Xp, Xp, same=True)
        gp.set_data(np.concatenate((X0, X4c)), np.concatenate((Y0, np.zeros((len(X4c), 1)))))
        Kpost = gp.predict_covariance(Xp, clip_cov=clip_cov)
    finally:
        gp.set_data(X0, Y0)
    return 0.5 * (np.linalg.slogdet(Kprior)[1] - np.linalg.slogdet(Kpost)[1])


def label_fidelity(var, fidlevs, open_top=True):
    """``GraceRIGV3.py:529-533`` (open_top) / ``:508-512`` (bounded l3): fidelity index of a
    candidate point from its localisation variance; strict inequalities, ties -> 0."""
    var = np.asarray(var, float)
    l1 = var < fidlevs[0]
    l2 = (var > fidlevs[0]) & (var < fidlevs[1])
    return (l1 * 2 + l2 * 1).astype(int)


def weighted_mse(err, cov):
    """``GPTrainers.py:121-137``: e^T (S^-1/||S^-1||_F) e / M."""
    inv = np.linalg.inv(cov)
    return float((err.T @ (inv / np.linalg.norm(inv)) @ err).item()) / err.shape[0]
