#!/usr/bin/env python
"""The reference's informationGainTest.py (lines 1-52) with `import GPy` swapped for the gpcore mirror: the
log-det information gain of a measurement set (on a test grid and on the measurements themselves) next to the
sequential sum 0.5 log(1 + sigma^2 / sigma_n).  Only the import line and the (commented-out) seed differ.

    python examples/information_gain_test.py [seed]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

entry.setup_path()
import gpcore.GPy as GPy  # noqa: E402            was: import GPy


def run(seed=0, measNum=3, verbose=True):
    np.random.seed(seed)                       # informationGainTest.py:5 (commented out there)
    Xpred = np.array([np.arange(-3, 3, .1)]).T
    X = np.random.uniform(-3., 3., (measNum, 1))
    Y = np.sin(X) + np.random.randn(measNum, 1) * 0.05
    priorX = np.array([[-100]])
    priorY = np.array([[0]])
    kernel = GPy.kern.RBF(input_dim=1, variance=0.7407926234918235, lengthscale=1.5704374366230516)
    m = GPy.models.GPRegression(X, Y, kernel)
    m.Gaussian_noise.variance = 0.0010413149736387451
    m.set_XY(priorX, priorY)
    _, Kprior = m.predict(Xpred, full_cov=1)
    _, Kprior2 = m.predict(X, full_cov=1)
    logDetPrior = np.log(np.linalg.det(Kprior))
    logDetPrior2 = np.log(np.linalg.det(Kprior2))
    m.set_XY(X, Y)
    _, Kposterior = m.predict(Xpred, full_cov=1)
    _, Kposterior2 = m.predict(X, full_cov=1)
    I = 0.5 * (logDetPrior - np.log(np.linalg.det(Kposterior)))        # I(f(Xpred); ftrue)
    I3 = 0.5 * (logDetPrior2 - np.log(np.linalg.det(Kposterior2)))     # I(f(X); ftrue)

    sig_n = m.Gaussian_noise.variance[0]
    xtemp = X[0, :]
    xtemp.shape = (1, X.shape[1])
    m.set_XY(xtemp, np.array([[0]]))
    _, sig_y = m.predict(xtemp)
    I2 = 0.5 * np.log(1 + sig_y[0, 0] / sig_n)
    for i in range(2, X.shape[0]):             # the reference skips index 1 (informationGainTest.py:45)
        xtemp = X[i, :]
        xtemp.shape = (1, X.shape[1])
        _, sig_y = m.predict(xtemp)
        I2 += 0.5 * np.log(1 + sig_y[0, 0] / sig_n)
        m.set_XY(np.concatenate((m.X, xtemp)), np.concatenate((m.Y, np.array([[0]]))))
        if verbose:
            print(m.X.shape, I2)
    if verbose:
        print(I, I2, I3)
    return I, I2, I3


if __name__ == "__main__":
    run(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
