"""Evaluation and I/O helpers of the offline trainer (``GPTrainers.py:25-58,115-165``), with the
dense algebra on the device: the covariance-weighted MSE goes through one Cholesky instead of
``np.linalg.inv`` of a 2000 x 2000 matrix (four of them per dataset in the reference).
"""
import numpy as np

from . import _lib as L
from .core import GPCore

_core = {}


def _handle(device=0):
    if device not in _core:
        _core[device] = GPCore(L.KIND_SF_RBF, 1, device)
    return _core[device]


def weighted_mse(err, cov, normalize=True, device=0):
    """``GPTrainers.py:121-137``: e^T (inv(S) / ||inv(S)||_F) e / M  (``normalize=False`` drops the
    Frobenius factor, the reference's ``Normalize=0`` branch).  Returns a float."""
    err = np.asarray(err, dtype=float).reshape(-1)
    quad, fro, _ = _handle(device).spd_stats(cov, err)
    return quad / (fro if normalize else 1.0) / err.shape[0]


def rmse(err):
    """``GPTrainers.py:141``."""
    return float(np.sqrt(np.mean(np.asarray(err, dtype=float) ** 2)))


GPDATA_HEADER = ["t", "x", "y", "z", "xh", "yh", "zh", "fieldVal", "fidLev"]   # prepGPData.py:48


def read_gpdata_csv(path, t_max=3600.0):
    """``GPTrainers.py:29-37``: header line + comma-separated rows, rows with t >= t_max dropped.
    Returns a dict of columns."""
    with open(path, "r") as f:
        headers = f.readline().strip().split(",")
        data = np.loadtxt(f, delimiter=",", ndmin=2)
    data = data[data[:, headers.index("t")] < t_max]
    return {h: data[:, i] for i, h in enumerate(headers)}


def split_fidelities(cols, estimated=True):
    """``GPTrainers.py:38-61``: per-fidelity input / target lists ordered lowest fidelity first
    (fidLev 3, 2, 1 -> emukit indices 0, 1, 2)."""
    keys = ("xh", "yh", "zh") if estimated else ("x", "y", "z")
    X = np.stack([cols[k] for k in keys], axis=1)
    y = cols["fieldVal"][:, None]
    xs, ys = [], []
    for lev in (3, 2, 1):
        m = cols["fidLev"] == lev
        xs.append(X[m])
        ys.append(y[m])
    return xs, ys


def write_gpres_csv(path, testPoints, fTrue, musf, sigsf_diag, mumf, sigmf):
    """``GPTrainers.py:138`` layout: x,y,z,trueField,sfMean,sfVar,mfMean,mfVar."""
    out = np.concatenate([np.asarray(a, float).reshape(len(testPoints), -1)
                          for a in (testPoints, fTrue, musf, sigsf_diag, mumf, sigmf)], axis=1)
    np.savetxt(path, out, delimiter=",", header=" x,y,z,trueField,sfMean,sfVar,mfMean,mfVar", comments="")


def write_mse_txt(path, rmse_by_model, wmse_by_model):
    """``GPTrainers.py:140-165``: ``RMSE <name>:<v>`` lines then ``WRMSE <name>:<v>`` lines."""
    with open(path, "w") as f:
        for k, v in rmse_by_model.items():
            f.write("RMSE {}:{}\n".format(k, v))
        for k, v in wmse_by_model.items():
            f.write("WRMSE {}:{}\n".format(k, v))


def read_mse_txt(path):
    out = {}
    with open(path) as f:
        for line in f:
            if ":" in line:
                k, v = line.strip().split(":", 1)
                out[k] = float(v.strip("[] "))
    return out
