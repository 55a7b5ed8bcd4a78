"""Expected-information-density map over the workspace grid: ``getEID`` of the reference
(``exploreSimSettings.py:6-37`` -- simulation -- and ``PhysicalExperimentCode/exploreExpSettings.py:8-30`` --
experiment), which turns one posterior mean / variance pass over the grid into the softmax of a UCB-like score
for the ergodic planner.  Only ``predict`` of the mirrored model runs on the device; the rest is O(G) host work."""
import numpy as np


def softmax(a):
    """``ergodicKLDivergence.py:6-9`` (no max-shift, like the reference)."""
    ea = np.exp(a)
    return ea / np.sum(ea)


def workspace_grid(WS, mD, counts=(10, 20, 10)):
    """The default test set of the simulation variant: ``np.meshgrid`` + C-order ``ravel`` (``:8-11``)."""
    WS = np.asarray(WS, dtype=float)
    specs = [[WS[0, 0], WS[0, 1], counts[0]], [WS[1, 0], WS[1, 1], counts[1]], [0, mD, counts[2]]]
    g = np.meshgrid(*[np.linspace(s[0], s[1], s[2]) for s in specs])
    return np.array([gi.ravel() for gi in g]).T


def getEID(gp, WS, mD, testSet=None, emu=False, alpha=1.0 / 11, auto=0, variant="sim", default_grid=None):
    """Returns ``(EID, ss3D)``.

    gp: mirrored ``GPRegression`` (``emu=False``) or ``GPyMultiOutputWrapper`` (``emu=True``: the grid is
    queried at fidelity index 2, the prior variance is ``param_array[[0, 4, 8, -1]].sum()``).
    variant "sim": ``fauxUCB = alpha mu + (1 - alpha) sqrt(|var|)`` and a uniform map when any variance is
    negative; variant "exp": negative variances are replaced by the prior variance first, default
    ``alpha = 0.2`` there, and the default test set is the module's ``ERGfieldGrid`` (pass ``default_grid``).
    auto: the module-level switch of the reference (``alpha = 1 - mean(var) / prior``).
    (The simulation variant of the reference tests ``testSet == None`` and therefore only accepts ``None``
    or a non-array; an array is accepted here.)"""
    if testSet is None:
        ss3D = workspace_grid(WS, mD) if variant == "sim" else np.asarray(default_grid, dtype=float)
    else:
        ss3D = np.asarray(testSet, dtype=float)
    if emu:
        mu, sig = gp.predict(np.hstack((ss3D, np.ones((ss3D.shape[0], 1)) * 2)))
        prior_sig = np.sum(np.asarray(gp.gpy_model.param_array)[[0, 4, 8, -1]])
    else:
        mu, sig = gp.predict(ss3D)
        prior_sig = gp.kern.variance[0] + gp.Gaussian_noise.variance[0]
    mu, sig = np.array(mu, dtype=float), np.array(sig, dtype=float)
    if variant == "exp":
        sig[sig < 0] = prior_sig
    if auto:
        alpha = 1 - np.mean(sig) / prior_sig
    if variant == "exp":
        return softmax(alpha * mu + (1 - alpha) * np.sqrt(sig)), ss3D
    EID = softmax(alpha * mu + (1 - alpha) * np.sqrt(np.abs(sig)))
    if np.any(sig < 0):
        EID = EID * 0 + 1 / EID.shape[0]
    return EID, ss3D
