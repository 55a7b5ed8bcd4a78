"""Drop-in for the reference module ``NIGP.py`` (import as ``from gpcore.nigp import NIGP``) (noisy-input GP, McHutchon & Rasmussen 2011):
same module functions, same ``NIGP`` class, same argument meaning and error behaviour -- every
dense step (kernel matrix, Cholesky, solves, log-det, posterior mean / gradients / variance /
covariance) runs in ``libgpcore.so`` on the GPU.  The host keeps only what the reference keeps
on the host: the alternation loop and SciPy's L-BFGS-B driver (``NIGP.py:215-243``).

Reference lines each function replaces are cited in the docstrings.
Restrictions: at most 3 input dimensions (the core's rows are (x, y, z, fidelity)).
"""
import numpy as np
from scipy.optimize import minimize

from . import _lib as L
from .core import GPCore, to_x4

_core_cache = {}


def _core(device=0):
    c = _core_cache.get(device)
    if c is None:
        c = _core_cache[device] = GPCore(L.KIND_NIGP, 1, device)
    return c


def _hyp5(lengthscales, sigma_f, sigma_y):
    ls = np.ones(3)
    l_in = np.asarray(lengthscales, dtype=float).ravel()
    ls[:l_in.size] = l_in
    return np.concatenate([ls, [float(sigma_f), float(sigma_y)]])


def SE_ARD_kernel(X1, X2, lengthscales, sigma_f, device=0):
    """``NIGP.py:11-20``: K = sigma_f * exp(-1/2 sum_d ((x_d - x'_d)/l_d)^2); ``sigma_f`` is the
    kernel *variance* (the reference hands it to ``GPy.kern.RBF(variance=sigma_f)``)."""
    c = _core(device)
    c.set_hypers(_hyp5(lengthscales, sigma_f, 1.0), 0.0)
    return c.kernel_matrix(to_x4(X1), to_x4(X2))


def compute_post_mean_and_gradients(X_train, y, lengthscales, sigma_f, sigma_y, noise_diag=None, device=0):
    """``NIGP.py:29-65``: posterior mean at the training inputs and its input gradients
    ``grads[i, d] = sum_j alpha_j K_ij (-(x_id - x_jd) / l_d^2)`` with
    ``alpha = (K + diag(sigma_y^2 + noise_diag))^-1 y`` (no jitter)."""
    X_train = np.asarray(X_train, dtype=float)
    N, D = X_train.shape
    c = _core(device)
    c.set_hypers(_hyp5(lengthscales, sigma_f, sigma_y), 0.0)
    X4 = to_x4(X_train)
    c.set_data(X4, np.asarray(y, dtype=float).ravel(), noise_diag)
    c.factor()
    f_mean, grads = c.mean_grad(X4)
    return f_mean, grads[:, :D]


def neg_log_marginal_likelihood(log_hyp, X, y, grad_fixed, noise_diag_extra_fixed=None, device=0):
    """``NIGP.py:130-165``: hypers ``[log l (D), log sigma_f, log sigma_y, log sigma_x (D)]``;
    per-point variance ``sigma_y^2 + sum_d grad_d^2 sigma_x_d^2 (+ extra)``, jitter 1e-8;
    returns 1e25 when the covariance is not positive definite."""
    X = np.asarray(X, dtype=float)
    N, D = X.shape
    log_hyp = np.asarray(log_hyp, dtype=float)
    ls = np.exp(log_hyp[:D])
    sigma_f = np.exp(log_hyp[D])
    sigma_y = np.exp(log_hyp[D + 1])
    sigma_x = np.exp(log_hyp[D + 2:])
    v = np.sum((np.asarray(grad_fixed) ** 2) * (sigma_x[None, :] ** 2), axis=1)
    if noise_diag_extra_fixed is not None:
        v = v + noise_diag_extra_fixed
    c = _core(device)
    try:
        c.set_hypers(_hyp5(ls, sigma_f, sigma_y), 1e-8)
        c.set_data(to_x4(X), np.asarray(y, dtype=float).ravel(), v)
        nlml, _ = c.factor()
    except (np.linalg.LinAlgError, ValueError):
        return 1e25
    return float(nlml)


def nlml_and_grad(log_hyp, X, y, grad_fixed, noise_diag_extra_fixed=None, device=0):
    """The objective of ``NIGP.py:130-165`` together with its analytic gradient with respect to
    the log hyper-parameters ``[log l (D), log sigma_f, log sigma_y, log sigma_x (D)]`` (the
    reference differentiates numerically inside L-BFGS-B).  Returns (1e25, zeros) when not PD."""
    X = np.asarray(X, dtype=float)
    N, D = X.shape
    log_hyp = np.asarray(log_hyp, dtype=float)
    hyp = np.exp(log_hyp)
    ls, sigma_f, sigma_y, sigma_x = hyp[:D], hyp[D], hyp[D + 1], hyp[D + 2:]
    G2 = np.asarray(grad_fixed) ** 2
    v = np.sum(G2 * (sigma_x[None, :] ** 2), axis=1)
    if noise_diag_extra_fixed is not None:
        v = v + noise_diag_extra_fixed
    c = _core(device)
    try:
        c.set_hypers(_hyp5(ls, sigma_f, sigma_y), 1e-8)
        c.set_data(to_x4(X), np.asarray(y, dtype=float).ravel(), v)
        nlml, _ = c.factor()
        g5, dW = c.nlml_grad(5, want_diag=True)
    except (np.linalg.LinAlgError, ValueError):
        return 1e25, np.zeros_like(log_hyp)
    if not np.isfinite(nlml):
        return 1e20, np.zeros_like(log_hyp)
    grad = np.concatenate([g5[:D] * ls, [g5[3] * sigma_f, g5[4] * sigma_y], (sigma_x ** 2) * (dW @ G2)])
    return float(nlml), grad


def safe_obj(lh, X, y, grad_fixed, noise_diag_extra_fixed):
    """``NIGP.py:119-123``."""
    val = neg_log_marginal_likelihood(lh, X, y, grad_fixed, noise_diag_extra_fixed)
    if not np.isfinite(val):
        return 1e20
    return val


def _median_pairwise(X):
    """``NIGP.py:200-202`` without the N x N x D temporary."""
    from scipy.spatial.distance import pdist
    d = pdist(X) if X.shape[0] > 1 else np.zeros(0)
    d = d[d > 0]
    return float(np.median(d)) if d.size else 1.0


class NIGP:
    """Same constructor, attributes and methods as the reference class (``NIGP.py:170-333``)."""

    def __init__(self, n_restarts=3, iters=3, verbose=True, device=0, analytic_grad=True):
        self.analytic_grad = analytic_grad
        self.n_restarts = n_restarts
        self.iters = iters
        self.verbose = verbose
        self.device = device
        self.lengthscales_ = None
        self.sigma_f_ = None
        self.sigma_y_ = None
        self.sigma_x_ = None
        self.X_train_ = None
        self.y_train_ = None
        self.noise_diag_train_ = None
        self._gp = None        # private core holding the factor used by predict
        self._stamp = None

    def get_params(self):
        return np.hstack((self.sigma_x_, self.sigma_f_, self.sigma_y_, self.lengthscales_))

    # -- fitting ------------------------------------------------------------------------------
    @staticmethod
    def _initial_log_hypers(X, y):
        """Starting point of ``NIGP.py:198-212``: median pairwise distance for every
        lengthscale, std(y) for sigma_f, a tenth of it for sigma_y, 1 % of std(X) for sigma_x."""
        D = X.shape[1]
        med = _median_pairwise(X)
        sy = np.std(y)
        sf0 = sy if sy > 0 else 1.0
        sx0 = np.maximum(0.01 * np.std(X, axis=0), 1e-8)
        return np.log(np.concatenate([np.full(D, med if med > 0 else 1.0), [sf0, 0.1 * sf0], sx0]))

    def _optimise(self, start, X, y, grads, maxiter_opt):
        """Step B (``NIGP.py:227-243``): ``n_restarts`` L-BFGS-B runs from ``start`` perturbed by
        0.1 * randn (NumPy's global generator, as in the reference, so ``np.random.seed``
        reproduces a run); box bounds [1e-6, 1e6] on every hyper-parameter."""
        box = [(np.log(1e-6), np.log(1e6))] * start.size
        zeros = np.zeros(X.shape[0])
        winner, last = None, None
        for _ in range(self.n_restarts):
            x0 = start + 0.1 * np.random.randn(*start.shape)
            if self.analytic_grad:   # device gradient: one factorisation (+ half) per L-BFGS-B evaluation
                last = minimize(nlml_and_grad, x0, args=(X, y, grads, zeros, self.device), jac=True,
                                method="L-BFGS-B", bounds=box, options={"maxiter": maxiter_opt})
            else:                    # the reference's own scheme: SciPy differentiates numerically
                last = minimize(safe_obj, x0, args=(X, y, grads, zeros), method="L-BFGS-B", bounds=box,
                                options={"maxiter": maxiter_opt})
            if last.fun < (1e99 if winner is None else winner.fun):
                winner = last
        chosen = last if winner is None else winner
        return chosen.x, (1e99 if winner is None else winner.fun)

    def fit(self, X, y, maxiter_opt=200):
        """``NIGP.py:191-260``: alternate (A) posterior-mean input gradients under the current
        hypers with the input-noise term switched off (``noise_diag=None``, ``:222``) and (B) NLML
        optimisation with those gradients frozen.  The stored ``noise_diag_train_`` uses the
        gradients of the last round (``:251-252``)."""
        X = np.asarray(X, dtype=float)
        y = np.asarray(y, dtype=float).flatten()
        D = X.shape[1]
        self.X_train_, self.y_train_ = X, y
        log_hyp = self._initial_log_hypers(X, y)
        grads = np.zeros_like(X)
        for it in range(self.iters):
            if self.verbose:
                print(f"NIGP iteration {it+1}/{self.iters} ...")
            _, grads = compute_post_mean_and_gradients(X, y, np.exp(log_hyp[:D]), np.exp(log_hyp[D]),
                                                       np.exp(log_hyp[D + 1]), device=self.device)
            log_hyp, val = self._optimise(log_hyp, X, y, grads, maxiter_opt)
            if self.verbose:
                print(f"  optimized nlml: {val:.6g}")
        hyp = np.exp(log_hyp)
        self.lengthscales_, self.sigma_f_, self.sigma_y_, self.sigma_x_ = hyp[:D], hyp[D], hyp[D + 1], hyp[D + 2:]
        self.noise_diag_train_ = np.sum((grads ** 2) * (self.sigma_x_[None, :] ** 2), axis=1)
        if self.verbose:
            print("Learned hyperparameters:")
            print(" lengthscales:", self.lengthscales_)
            print(" sigma_f:", self.sigma_f_)
            print(" sigma_y:", self.sigma_y_)
            print(" sigma_x (per-dim):", self.sigma_x_)
        return self

    # -- factor cache (the reference re-assembles and re-factors on every predict call,
    #    NIGP.py:284-289; here the factor is kept until a public attribute changes) -------------
    def _factor(self):
        X = np.asarray(self.X_train_, dtype=float)
        y = np.asarray(self.y_train_, dtype=float).ravel()
        nd = self.noise_diag_train_
        hyp = _hyp5(self.lengthscales_, self.sigma_f_, self.sigma_y_)
        stamp = (hyp.tobytes(), X.shape, hash(X.tobytes()), hash(y.tobytes()),
                 None if nd is None else hash(np.asarray(nd, float).tobytes()))
        if self._gp is None:
            self._gp = GPCore(L.KIND_NIGP, 1, self.device)
        if stamp != self._stamp:
            self._gp.set_hypers(hyp, 0.0)          # NIGP.predict adds no jitter (NIGP.py:287-288)
            self._gp.set_data(to_x4(X), y, nd)
            self._gp.factor()
            self._stamp = stamp
        return self._gp

    def predict_grid_mean(self, ax, ay, az):
        """Extension: ``predict(grid, return_var=False)`` for the tensor grid ``np.meshgrid(ax, ay, az, indexing="ij")``
        as FP64 tensor-core GEMMs; shape ``(len(ax), len(ay), len(az))``."""
        return self._factor().predict_grid_mean(ax, ay, az, 0)

    def predict(self, Xs, Xs_input_noise=None, return_var=True, return_cov=False):
        """``NIGP.py:269-333``.  ``return_cov=False`` computes only the diagonal (the reference
        always builds the M x M matrix, ``:299``), with identical values."""
        Xs = np.asarray(Xs, dtype=float)
        M, D = Xs.shape
        gp = self._factor()
        Xs4 = to_x4(Xs)
        if not (return_var or return_cov):
            mean, _ = gp.predict(Xs4, 0, want_var=False)
            return mean
        sx3 = None
        if Xs_input_noise is not None:
            sx = np.asarray(Xs_input_noise)
            if sx.ndim == 1 and sx.size == D:
                sx3 = np.zeros((1, 3))
                sx3[0, :D] = sx
            elif sx.shape == (M, D):
                sx3 = np.zeros((M, 3))
                sx3[:, :D] = sx
            else:
                raise ValueError("Xs_input_noise must be scalar, shape (D,) or (M,D)")
        if return_cov:
            extra = None
            if sx3 is not None:
                _, grads = gp.mean_grad(Xs4)
                extra = np.sum(grads ** 2 * sx3 ** 2, axis=1)
            mean, cov = gp.predict_cov(Xs4, L.NIGP_FLOOR, extra_diag=extra)
            return mean, cov
        if sx3 is None:
            return gp.predict(Xs4, L.NIGP_FLOOR)
        mean, var = np.empty(M), np.empty(M)
        sx3 = np.ascontiguousarray(sx3)
        gp._ck(gp.lib.gpc_predict_noisy(gp.h, L.dptr(Xs4), M, L.dptr(sx3), sx3.shape[0], L.dptr(mean), L.dptr(var),
                                        L.NIGP_FLOOR))
        return mean, var
