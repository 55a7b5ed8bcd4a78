"""``GPy.kern.RBF`` / ``GPy.kern.Matern32`` restated (stationary kernels, ARD).

GPy ``Stationary._scaled_dist`` / ``_unscaled_dist``: inputs are divided by the
lengthscale, the squared distance is formed with the Gram identity
``|a|^2 + |b|^2 - 2 a.b``, its diagonal forced to zero when X2 is None, clipped at
0 and square-rooted;  ``RBF.K_of_r = variance * exp(-0.5 r^2)``,
``Matern32.K_of_r = variance * (1 + sqrt(3) r) * exp(-sqrt(3) r)``.
With ``inv_l=True`` GPy stores 1/l^2 and recovers l = sqrt(1/inv_l).
"""
import numpy as np


def _scaled_dist(X, X2, lengthscale):
    A = X / lengthscale
    if X2 is None:
        sq = np.sum(np.square(A), 1)
        r2 = -2.0 * A.dot(A.T) + (sq[:, None] + sq[None, :])
        r2[np.diag_indices(A.shape[0])] = 0.0
    else:
        B = X2 / lengthscale
        r2 = -2.0 * A.dot(B.T) + (np.sum(np.square(A), 1)[:, None] + np.sum(np.square(B), 1)[None, :])
    return np.sqrt(np.clip(r2, 0.0, np.inf))


class _Stationary:
    def __init__(self, input_dim, variance=1.0, lengthscale=None, ARD=False, inv_l=False):
        self.input_dim = int(input_dim)
        self.variance = np.atleast_1d(np.asarray(variance, dtype=float))
        if lengthscale is None:
            lengthscale = np.ones(self.input_dim if ARD else 1)
        ls = np.atleast_1d(np.asarray(lengthscale, dtype=float))
        if inv_l:
            ls = np.sqrt(1.0 / (1.0 / ls ** 2))
        self.lengthscale = ls
        self.ARD = ARD

    def K(self, X, X2=None):
        X = np.asarray(X, dtype=float)
        X2 = None if X2 is None else np.asarray(X2, dtype=float)
        return self.K_of_r(_scaled_dist(X, X2, self.lengthscale))

    def Kdiag(self, X):
        return np.full(np.asarray(X).shape[0], float(self.variance[0]))


class RBF(_Stationary):
    def K_of_r(self, r):
        return self.variance[0] * np.exp(-0.5 * r ** 2)


class Matern32(_Stationary):
    def K_of_r(self, r):
        return self.variance[0] * (1.0 + np.sqrt(3.0) * r) * np.exp(-np.sqrt(3.0) * r)
