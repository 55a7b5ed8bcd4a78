"""Host-side mirror of the GP model protocol the reference consumes (GPy ``GPRegression``,
emukit ``GPyLinearMultiFidelityModel`` + ``GPyMultiOutputWrapper``), backed by ``libgpcore.so``.

Only the surface the reference's trainers / planners touch is mirrored:

    GPTrainers.py:62-69,80-98,116-120            fit + predict + predict_covariance
    PhysicalExperimentCode/GraceRIGV3.py:446-678  copy / set_XY / set_data / predict inside IG
    exploreSimSettings.py:15-19                   param_array[[0,4,8,-1]], kern.variance[0]
    PhysicalExperimentCode/...MFGP.py:386-420     param_array slice assignment, optimize()

State lives in ONE flat ``param_array`` in the reference's order; ``kern.variance``,
``kern.lengthscale``, ``Gaussian_noise.variance`` ... are views into it, so slice assignment
on ``param_array`` and attribute assignment on the parts stay coherent.  The device factor is
rebuilt lazily whenever (parameters, data) changed since the last factorisation.
"""
import copy as _copy
import json

import numpy as np
from scipy.optimize import minimize

from . import _lib as L
from .core import GPCore, to_x4

GPY_JITTER = 1e-8  # GPy exact_gaussian_inference: K + (noise + 1e-8) I


# ----------------------------------------------------------------------------------------------
# parameters
# ----------------------------------------------------------------------------------------------
class Param(np.ndarray):
    """ndarray view with the paramz methods the reference calls (``fix``, ``constrain_*``)."""

    def __new__(cls, values, name="param"):
        obj = np.atleast_1d(np.asarray(values, dtype=float)).copy().view(cls)
        obj._meta = {"name": name, "fixed": False, "bounds": None}
        return obj

    def __array_finalize__(self, obj):
        self._meta = getattr(obj, "_meta", {"name": "param", "fixed": False, "bounds": None})

    def __deepcopy__(self, memo):
        new = Param(np.asarray(self), self._meta["name"])
        new._meta = dict(self._meta)
        return new

    @classmethod
    def view_of(cls, buf, meta):
        v = buf.view(cls)
        v._meta = meta
        return v

    def fix(self, value=None):
        if value is not None:
            self[...] = value
        self._meta["fixed"] = True

    def unfix(self):
        self._meta["fixed"] = False

    def constrain_bounded(self, lower, upper):
        self._meta["bounds"] = (float(lower), float(upper))

    def constrain_positive(self):
        self._meta["bounds"] = None

    constrain_fixed = fix
    unconstrain_fixed = unfix


class _Parameterized:
    """Owns an ordered list of (attribute name, Param) and can re-bind them onto a flat buffer."""

    def _param_items(self):
        raise NotImplementedError

    def _bind(self, flat, offset, prefix):
        names = []
        for attr, p in self._param_items():
            n = p.size
            flat[offset:offset + n] = np.asarray(p)
            object.__setattr__(self, "_p_" + attr, Param.view_of(flat[offset:offset + n], p._meta))
            names.append((prefix + attr, n))
            offset += n
        return offset, names


def _param_property(attr):
    def get(self):
        return getattr(self, "_p_" + attr)

    def set_(self, value):
        getattr(self, "_p_" + attr)[...] = value

    return property(get, set_)


# ----------------------------------------------------------------------------------------------
# kernels (GPy.kern.RBF / GPy.kern.Matern32) and likelihoods
# ----------------------------------------------------------------------------------------------
class _Stationary(_Parameterized):
    _base = 0
    name = "stationary"

    def __init__(self, input_dim, variance=1.0, lengthscale=None, ARD=False, inv_l=False, name=None):
        self.input_dim = int(input_dim)
        self.ARD = bool(ARD)
        if self.input_dim > 3:
            raise ValueError("gpcore kernels support at most 3 input dimensions")
        if lengthscale is None:
            lengthscale = np.ones(self.input_dim if ARD else 1)
        ls = np.atleast_1d(np.asarray(lengthscale, dtype=float))
        if ARD and ls.size == 1:
            ls = np.repeat(ls, self.input_dim)
        if ls.size not in (1, self.input_dim):
            raise ValueError("lengthscale must have 1 or input_dim entries")
        if name:
            self.name = name
        self._p_variance = Param(variance, "variance")
        self._p_lengthscale = Param(ls, "lengthscale")

    variance = _param_property("variance")
    lengthscale = _param_property("lengthscale")

    def _param_items(self):
        return [("variance", self._p_variance), ("lengthscale", self._p_lengthscale)]

    def ls3(self):
        ls = np.asarray(self._p_lengthscale, dtype=float)
        out = np.ones(3)
        out[:self.input_dim] = ls if ls.size > 1 else ls[0]
        return out

    def _sf_hypers(self, noise=1.0):
        return np.concatenate([[float(self._p_variance[0])], self.ls3(), [noise]])

    def K(self, X, X2=None):
        """``kern.K`` evaluated on the device (no noise)."""
        core = _kernel_core(L.KIND_SF_MAT32 if self._base else L.KIND_SF_RBF, 1, 0)
        core.set_hypers(self._sf_hypers(), 0.0)
        return core.kernel_matrix(to_x4(X), None if X2 is None else to_x4(X2))

    def Kdiag(self, X):
        return np.full(np.asarray(X).shape[0], float(self._p_variance[0]))

    def copy(self):
        return _copy.deepcopy(self)


class RBF(_Stationary):
    _base = 0
    name = "rbf"


class Matern32(_Stationary):
    _base = 1
    name = "Mat32"


class Gaussian(_Parameterized):
    """``GPy.likelihoods.Gaussian``."""
    name = "Gaussian_noise"

    def __init__(self, variance=1.0, name=None):
        if name:
            self.name = name
        self._p_variance = Param(variance, "variance")

    variance = _param_property("variance")

    def _param_items(self):
        return [("variance", self._p_variance)]


class MixedNoise(_Parameterized):
    """``GPy.likelihoods.MixedNoise`` over per-fidelity Gaussians (emukit default likelihood)."""
    name = "mixed_noise"

    def __init__(self, likelihoods_list):
        self.likelihoods_list = list(likelihoods_list)
        for i, lk in enumerate(self.likelihoods_list):
            lk.name = "Gaussian_noise" if i == 0 else "Gaussian_noise_%d" % i
            setattr(self, lk.name, lk)

    def _bind(self, flat, offset, prefix):
        names = []
        for lk in self.likelihoods_list:
            offset, nm = lk._bind(flat, offset, prefix + lk.name + ".")
            names += nm
        return offset, names


class LinearMultiFidelityKernel(_Parameterized):
    """``emukit.multi_fidelity.kernels.LinearMultiFidelityKernel``: Kennedy-O'Hagan AR1 sum
    over per-fidelity stationary kernels, ``scale`` = rho_1 .. rho_{F-1} (initially 1)."""
    name = "multifidelity"

    def __init__(self, kernels):
        self.kernels = list(kernels)
        self.n_fidelities = len(self.kernels)
        if not 1 <= self.n_fidelities <= 4:
            raise ValueError("1..4 fidelities supported")
        base = {k._base for k in self.kernels}
        if len(base) != 1:
            raise ValueError("all per-fidelity kernels must be of the same family")
        self._base = base.pop()
        counts = {}
        for k in self.kernels:  # GPy names duplicates rbf, rbf_1, rbf_2 ...
            i = counts.get(k.name, 0)
            counts[k.name] = i + 1
            attr = k.name if i == 0 else "%s_%d" % (k.name, i)
            k._attr = attr
            setattr(self, attr, k)
        self._p_scale = Param(np.ones(self.n_fidelities - 1), "scale")

    scale = _param_property("scale")

    def _bind(self, flat, offset, prefix):
        names = []
        for k in self.kernels:
            offset, nm = k._bind(flat, offset, prefix + k._attr + ".")
            names += nm
        n = self._p_scale.size
        flat[offset:offset + n] = np.asarray(self._p_scale)
        self._p_scale = Param.view_of(flat[offset:offset + n], self._p_scale._meta)
        names.append((prefix + "scale", n))
        return offset + n, names

    def _mf_hypers(self, noise):
        parts = []
        for k in self.kernels:
            parts += [[float(k._p_variance[0])], k.ls3()]
        parts += [np.asarray(self._p_scale, float), np.atleast_1d(noise)]
        return np.concatenate(parts)

    def K(self, X, X2=None):
        """``gpy_model.kern.K(X4)`` (``GraceRIGV3.py:515``): rows carry the fidelity index last."""
        F = self.n_fidelities
        core = _kernel_core(L.KIND_MF_AR1_MAT32 if self._base else L.KIND_MF_AR1_RBF, F, 0)
        core.set_hypers(self._mf_hypers(np.ones(1)), 0.0)
        return core.kernel_matrix(_x4_mf(X, F), None if X2 is None else _x4_mf(X2, F))


_KERNEL_CORES = {}


def _kernel_core(kind, F, device):
    """One cached handle per (kind, F, device) for ``kern.K`` calls (a handle owns three streams and ~17 events:
    creating and destroying one per call costs more than the kernel matrix of a planner-sized input)."""
    key = (int(kind), int(F), int(device))
    core = _KERNEL_CORES.get(key)
    if core is None or getattr(core, "h", None) is None:
        core = _KERNEL_CORES[key] = GPCore(kind, F, device)
    return core


_HOST_CHECK_ROWS = 1 << 16   # larger inputs are validated on the device (gpc_predict: GpcoreArgError, a ValueError)


def _x4_mf(X, n_fidelities=None):
    """(n, D + 1) rows with the fidelity index last -> (n, 4) rows; 3-D inputs already have the
    device row layout (x, y, z, fid) and pass through without a copy.  With ``n_fidelities`` the labels are
    validated the way emukit does (integers in [0, F)) -- ``ValueError`` otherwise; the C ABI re-checks."""
    X = np.asarray(X, dtype=float)
    if X.ndim != 2 or X.shape[1] < 2:
        raise ValueError("multi-fidelity inputs are (n, D + 1) rows with the fidelity index in the last column")
    if n_fidelities is not None and 0 < X.shape[0] <= _HOST_CHECK_ROWS:
        f = X[:, -1]
        if not (np.all(f >= 0) and np.all(f < n_fidelities) and np.all(f == np.floor(f))):
            raise ValueError("fidelity index (last input column) must be an integer in [0, %d)" % n_fidelities)
    if X.shape[1] == 4 and X.flags["C_CONTIGUOUS"]:
        return X
    return to_x4(X[:, :-1], X[:, -1])


# ----------------------------------------------------------------------------------------------
# models
# ----------------------------------------------------------------------------------------------
def _softplus_inv(v):
    v = np.asarray(v, float)
    return np.where(v > 30, v, np.log(np.expm1(np.minimum(v, 30))))


def _softplus(x):
    return np.where(x > 30, x, np.log1p(np.exp(np.minimum(x, 30))))


def _transforms(lo, hi):
    """GPy's parameter transforms for a vector of free parameters: Logexp (softplus) where ``hi`` is infinite,
    Logistic for ``constrain_bounded(lo, hi)``:
        theta = lo + (hi - lo) / (1 + exp(-x)),  d theta / dx = (theta - lo)(hi - theta) / (hi - lo)
    -- smooth and strictly inside the bounds, so the optimiser never sees a flat objective with a non-zero gradient.
    Returns (from_raw, dtheta_dx, to_raw)."""
    lo, hi = np.asarray(lo, float), np.asarray(hi, float)
    bounded = np.isfinite(hi)
    span = hi[bounded] - lo[bounded]

    def from_raw(x):
        x = np.asarray(x, float)
        th = np.empty_like(x)
        th[~bounded] = _softplus(x[~bounded])
        th[bounded] = lo[bounded] + span / (1.0 + np.exp(-np.clip(x[bounded], -700.0, 700.0)))
        return th

    def dtheta_dx(th):
        d = np.empty_like(th)
        d[~bounded] = -np.expm1(-th[~bounded])                       # d softplus / dx = 1 - exp(-theta)
        d[bounded] = (th[bounded] - lo[bounded]) * (hi[bounded] - th[bounded]) / span
        return d

    def to_raw(th):
        th = np.asarray(th, float)
        x = np.empty_like(th)
        x[~bounded] = _softplus_inv(np.maximum(th[~bounded], 1e-12))
        u = np.clip((th[bounded] - lo[bounded]) / span, 1e-12, 1.0 - 1e-12)   # a start on / outside a bound moves inside
        x[bounded] = np.log(u) - np.log1p(-u)
        return x

    return from_raw, dtheta_dx, to_raw


class _DeviceGP:
    """Shared machinery: flat ``param_array``, lazy device factor, predict, optimise."""

    def _finish_init(self, device):
        self.device = device
        parts = [(self.kern, self.kern.name + "."), (self.likelihood, self.likelihood.name + ".")]
        n = sum(p.size for obj, _ in parts for p in self._all_params(obj))
        self.param_array = np.zeros(n)
        off, self._names = 0, []
        for obj, prefix in parts:
            off, nm = obj._bind(self.param_array, off, prefix)
            self._names += nm
        self._core = None
        self._stamp = None
        self._data_version = 0
        self.nlml_ = None

    @staticmethod
    def _all_params(obj):
        if isinstance(obj, LinearMultiFidelityKernel):
            return [p for k in obj.kernels for _, p in k._param_items()] + [obj._p_scale]
        if isinstance(obj, MixedNoise):
            return [p for lk in obj.likelihoods_list for _, p in lk._param_items()]
        return [p for _, p in obj._param_items()]

    def parameter_names(self):
        return [n for n, _ in self._names]

    def __getstate__(self):  # deep copies / pickles never carry the device handle
        d = self.__dict__.copy()
        d["_core"] = None
        d["_stamp"] = None
        return d

    def copy(self):
        new = _copy.deepcopy(self)
        # deepcopy breaks the view relationship: re-bind the parts onto the copied flat array
        flat = new.param_array.copy()
        new._finish_init(self.device)
        new.param_array[:] = flat
        new._data_version = 0
        return new

    # -- device state ---------------------------------------------------------------------
    def predict_grid_mean(self, ax, ay, az, fid=0):
        """Extension: the posterior mean on the tensor grid ``np.meshgrid(ax, ay, az, indexing="ij")`` (every test
        set of the reference is one) as FP64 tensor-core GEMMs -- shape ``(len(ax), len(ay), len(az))``; ``fid`` is the
        fidelity index of a multi-fidelity model.  Squared-exponential kernels only."""
        return self._ensure_factor().predict_grid_mean(ax, ay, az, fid)

    def _ensure_factor(self):
        stamp = (self.param_array.tobytes(), self._data_version)
        if self._core is None:
            self._core = GPCore(self._kind(), self._F(), self.device)
            self._stamp = None
        if stamp != self._stamp:
            self._core.set_hypers(self._flat_hypers(), GPY_JITTER)
            self._core.set_data(self._X4(), self.Y)
            self.nlml_, self.logdet_ = self._core.factor()
            self._stamp = stamp
        return self._core

    def objective_function(self):
        """GPy ``Model.objective_function``: negative log marginal likelihood."""
        self._ensure_factor()
        return self.nlml_

    def log_likelihood(self):
        return -self.objective_function()

    def objective_function_gradients(self):
        """d NLML / d param_array, analytic, on the device (``gpc_nlml_grad``)."""
        core = self._ensure_factor()
        return core.nlml_grad(self._flat_hypers().size)[self._grad_map()]

    def _grad_map(self):
        """Indices of the flat device hyper vector that correspond to ``param_array`` entries."""
        return np.arange(self.param_array.size)

    # -- optimisation (host L-BFGS-B over the device NLML; softplus-transformed positives) ----
    def _free_mask(self):
        mask = np.ones(self.param_array.size, bool)
        bounds = [None] * self.param_array.size
        off = 0
        for obj in (self.kern, self.likelihood):
            for p in self._all_params(obj):
                if p._meta["fixed"]:
                    mask[off:off + p.size] = False
                for i in range(p.size):
                    bounds[off + i] = p._meta["bounds"]
                off += p.size
        return mask, bounds

    def optimize(self, max_iters=1000, messages=False, analytic_gradients=True, **_):
        """``model.optimize()`` (``GPTrainers.py:68,84,94``): L-BFGS-B on the NLML, positive
        parameters through GPy's default Logexp (softplus) transform, ``constrain_bounded`` ones through
        its Logistic transform (``...MFGP.py:408-410,664-666``), fixed ones left alone.
        Every evaluation is one device assembly + Cholesky (+ the analytic gradient, about half a
        factorisation more); ``analytic_gradients=False`` falls back to forward differences."""
        mask, bounds = self._free_mask()
        idx = np.nonzero(mask)[0]
        if idx.size == 0:
            return self
        start = self.param_array.copy()

        lo = np.array([bounds[i][0] if bounds[i] is not None else 0.0 for i in idx], float)
        hi = np.array([bounds[i][1] if bounds[i] is not None else np.inf for i in idx], float)
        from_raw, dtheta_dx, to_raw = _transforms(lo, hi)

        def f(x):
            th = from_raw(x)
            self.param_array[idx] = th
            try:
                v = self.objective_function()
                if not np.isfinite(v):
                    raise np.linalg.LinAlgError("non-finite objective")
                if not analytic_gradients:
                    return v
                g = self.objective_function_gradients()[idx] * dtheta_dx(th)
            except np.linalg.LinAlgError:
                return (1e25, np.zeros(idx.size)) if analytic_gradients else 1e25
            return v, g

        x0 = to_raw(start[idx])
        best_x = x0
        first = f(x0)
        best_f = first[0] if analytic_gradients else first
        try:
            opts = {"maxiter": int(max_iters)}
            if not analytic_gradients:
                opts["eps"] = 1e-6
            res = minimize(f, x0, jac=bool(analytic_gradients), method="L-BFGS-B", options=opts)
            if res.fun < best_f:
                best_x, best_f = res.x, res.fun
        finally:
            self.param_array[idx] = from_raw(best_x)
        self._ensure_factor()
        return self

    def optimize_restarts(self, num_restarts=1, robust=True, **kw):
        best, best_f = None, np.inf
        for r in range(max(int(num_restarts), 1)):
            if r > 0:
                mask, _ = self._free_mask()
                self.param_array[mask] = np.abs(np.random.randn(int(mask.sum()))) + 1e-3
            try:
                self.optimize(**kw)
                fval = self.objective_function()
            except Exception:
                if not robust:
                    raise
                continue
            if fval < best_f:
                best, best_f = self.param_array.copy(), fval
        if best is not None:
            self.param_array[:] = best
        return self

    def _save_model(self, path, compress=False, save_data=True):
        """``gp._save_model`` (``...SFGP.py:386``): JSON dump of names, parameters and data."""
        d = {"class": type(self).__name__, "parameter_names": self.parameter_names(),
             "param_array": self.param_array.tolist()}
        if save_data:
            d["X"], d["Y"] = np.asarray(self.X).tolist(), np.asarray(self.Y).tolist()
        with open(path + ".json", "w") as fh:
            json.dump(d, fh)


class GPRegression(_DeviceGP):
    """``GPy.models.GPRegression(X, Y, kernel, noise_var=1.)`` -- zero mean, Gaussian likelihood.
    ``param_array = [variance, lengthscale(D | 1), noise_var]``."""

    def __init__(self, X, Y, kernel=None, Y_metadata=None, normalizer=None, noise_var=1.0, mean_function=None, device=0):
        # GPy's signature; the reference only ever passes mean_function=None (HowManyPoints.py:91-92)
        if mean_function is not None or normalizer not in (None, False):
            raise NotImplementedError("gpcore mirrors the zero-mean, un-normalised GPRegression the reference uses")
        X = np.asarray(X, dtype=float)
        if kernel is None:
            kernel = RBF(X.shape[1])
        self.kern = kernel
        self.likelihood = Gaussian(noise_var)
        self._finish_init(device)
        self.set_XY(X, Y)

    @property
    def Gaussian_noise(self):
        return self.likelihood

    def _kind(self):
        return L.KIND_SF_MAT32 if self.kern._base else L.KIND_SF_RBF

    def _F(self):
        return 1

    def _flat_hypers(self):
        return self.kern._sf_hypers(float(self.likelihood._p_variance[0]))

    def _X4(self):
        return to_x4(self.X)

    def objective_function_gradients(self):
        g = self._ensure_factor().nlml_grad(5)          # [variance, lx, ly, lz, noise]
        D, nl = self.kern.input_dim, self.kern._p_lengthscale.size
        gl = g[1:1 + D] if nl > 1 else np.array([np.sum(g[1:1 + D])])   # one shared lengthscale: chain rule sums
        return np.concatenate([[g[0]], gl, [g[4]]])

    def set_XY(self, X=None, Y=None):
        if X is not None:
            self.X = np.asarray(X, dtype=float)
        if Y is not None:
            self.Y = np.asarray(Y, dtype=float).reshape(-1, 1)
        if self.X.shape[0] != self.Y.shape[0]:
            raise ValueError("X and Y disagree on N")
        self._data_version += 1

    def predict(self, Xnew, full_cov=False, include_likelihood=True, **_):
        """GPy ``predict``: (mu (M,1), var (M,1) | (M,M)); the likelihood variance is included by
        default; the latent marginal variance is clipped at 1e-15 on the diagonal path."""
        core = self._ensure_factor()
        Xs4 = to_x4(Xnew)
        noise = L.INCLUDE_NOISE if include_likelihood else 0
        if full_cov:
            mu, cov = core.predict_cov(Xs4, noise)
            return mu[:, None], cov
        mu, var = core.predict(Xs4, noise | L.CLIP_DIAG)
        return mu[:, None], var[:, None]


class GPyLinearMultiFidelityModel(_DeviceGP):
    """``emukit.multi_fidelity.models.GPyLinearMultiFidelityModel(X, Y, kernel, n_fidelities,
    likelihood=None)``; X carries the fidelity index in its last column (0 = lowest)."""

    def __init__(self, X, Y, kernel, n_fidelities, likelihood=None, device=0):
        if not isinstance(kernel, LinearMultiFidelityKernel):
            raise TypeError("kernel must be a LinearMultiFidelityKernel")
        if kernel.n_fidelities != n_fidelities:
            raise ValueError("kernel and n_fidelities disagree")
        self.n_fidelities = int(n_fidelities)
        self.kern = kernel
        self.likelihood = likelihood if likelihood is not None else MixedNoise(
            [Gaussian(1.0) for _ in range(self.n_fidelities)])
        self._finish_init(device)
        self.set_XY(X, Y)

    def _kind(self):
        return L.KIND_MF_AR1_MAT32 if self.kern._base else L.KIND_MF_AR1_RBF

    def _F(self):
        return self.n_fidelities

    def _noise_vec(self):
        if isinstance(self.likelihood, MixedNoise):
            return np.array([float(lk._p_variance[0]) for lk in self.likelihood.likelihoods_list])
        return np.array([float(self.likelihood._p_variance[0])])

    def _flat_hypers(self):
        return self.kern._mf_hypers(self._noise_vec())

    def _X4(self):
        return _x4_mf(self.X, self.n_fidelities)

    def objective_function_gradients(self):
        F = self.n_fidelities
        n_dev = 4 * F + (F - 1) + self._noise_vec().size
        g = self._ensure_factor().nlml_grad(n_dev)
        out = []
        for m, k in enumerate(self.kern.kernels):       # (variance, lengthscale(D | 1)) per fidelity
            D, nl = k.input_dim, k._p_lengthscale.size
            gl = g[4 * m + 1:4 * m + 1 + D]
            out += [[g[4 * m]], gl if nl > 1 else [np.sum(gl)]]
        out += [g[4 * F:4 * F + F - 1], g[4 * F + F - 1:]]
        return np.concatenate([np.atleast_1d(o) for o in out])

    def set_XY(self, X=None, Y=None):
        if X is not None:
            X = np.asarray(X, dtype=float)
            if X.ndim != 2 or X.shape[1] < 2:
                raise ValueError("X must be (N, D + 1) with the fidelity index last")
            f = X[:, -1]
            if np.any(f < 0) or np.any(f >= self.n_fidelities) or np.any(f != np.floor(f)):
                raise ValueError("fidelity indices must be integers in [0, n_fidelities)")
            self.X = X
        if Y is not None:
            self.Y = np.asarray(Y, dtype=float).reshape(-1, 1)
        if self.X.shape[0] != self.Y.shape[0]:
            raise ValueError("X and Y disagree on N")
        self._data_version += 1

    def predict(self, Xnew, full_cov=False, include_likelihood=True, Y_metadata=None, **_):
        core = self._ensure_factor()
        Xs4 = _x4_mf(Xnew, self.n_fidelities)
        noise = L.INCLUDE_NOISE if include_likelihood else 0
        if full_cov:
            mu, cov = core.predict_cov(Xs4, noise)
            return mu[:, None], cov
        mu, var = core.predict(Xs4, noise | L.CLIP_DIAG)
        return mu[:, None], var[:, None]


class GPyMultiOutputWrapper:
    """``emukit.model_wrappers.gpy_model_wrappers.GPyMultiOutputWrapper``."""

    def __init__(self, gpy_model, n_outputs, n_optimization_restarts, verbose_optimization=True):
        self.gpy_model = gpy_model
        self.n_outputs = n_outputs
        self.n_optimization_restarts = n_optimization_restarts
        self.verbose_optimization = verbose_optimization

    @property
    def X(self):
        return self.gpy_model.X

    @property
    def Y(self):
        return self.gpy_model.Y

    def set_data(self, X, Y):
        self.gpy_model.set_XY(X, Y)

    def predict_grid_mean(self, ax, ay, az, fid=0):
        return self.gpy_model.predict_grid_mean(ax, ay, az, fid)

    def predict(self, X):
        """(mean, variance) at fidelity-indexed rows, per-fidelity noise included."""
        return self.gpy_model.predict(X)

    def predict_covariance(self, X, with_noise=True):
        """Full posterior covariance, clipped element-wise at 1e-10 like the emukit wrapper."""
        core = self.gpy_model._ensure_factor()
        flags = (L.INCLUDE_NOISE if with_noise else 0) | L.CLIP_COV
        _, cov = core.predict_cov(_x4_mf(X, self.gpy_model.n_fidelities), flags, want_mean=False)
        return cov

    def optimize(self):
        self.gpy_model.optimize_restarts(self.n_optimization_restarts, robust=True)

    def copy(self):
        return GPyMultiOutputWrapper(self.gpy_model.copy(), self.n_outputs, self.n_optimization_restarts,
                                     self.verbose_optimization)


# ----------------------------------------------------------------------------------------------
# emukit.multi_fidelity.convert_lists_to_array
# ----------------------------------------------------------------------------------------------
def convert_x_list_to_array(x_list):
    """Stack per-fidelity inputs and append the list index as the fidelity column."""
    if not all(np.asarray(x).ndim == 2 for x in x_list):
        raise ValueError("All x arrays must have 2 dimensions")
    return np.concatenate([np.hstack([np.asarray(x, float), np.full((len(x), 1), float(i))])
                           for i, x in enumerate(x_list)], axis=0)


def convert_y_list_to_array(y_list):
    if not all(np.asarray(y).ndim == 2 for y in y_list):
        raise ValueError("All y arrays must have 2 dimensions")
    return np.concatenate([np.asarray(y, float) for y in y_list], axis=0)


def convert_xy_lists_to_arrays(x_list, y_list):
    if len(x_list) != len(y_list):
        raise ValueError("Different number of fidelities between x and y")
    for x, y in zip(x_list, y_list):
        if len(x) != len(y):
            raise ValueError("Different number of points in x and y at one fidelity")
    return convert_x_list_to_array(x_list), convert_y_list_to_array(y_list)
