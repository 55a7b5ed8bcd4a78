"""``GPCore``: one handle of the CUDA inference core (assembly -> Cholesky -> alpha ->
posterior / information gain).  Thin, stateless-on-the-host wrapper over the C ABI; the model
classes in ``gpy_compat`` / ``emukit_compat`` / ``NIGP`` hold the reference-facing state.
"""
import ctypes as C

import numpy as np

from . import _lib as L


def to_x4(X, fid=None):
    """(n, D<=3) [+ fidelity column] -> C-contiguous (n, 4) rows (x, y, z, fidelity)."""
    X = np.asarray(X, dtype=np.float64)
    if X.ndim == 1:
        X = X[None, :]
    n, d = X.shape
    if d > 3:
        raise ValueError("gpcore supports at most 3 spatial input dimensions (got %d)" % d)
    out = np.zeros((n, 4), dtype=np.float64)
    out[:, :d] = X
    if fid is not None:
        out[:, 3] = fid
    return out


class GPCore:
    def __init__(self, kind, F=1, device=0):
        self.lib = L.load()
        self.kind, self.F, self.device = int(kind), int(F), int(device)
        h = C.c_void_p()
        rc = self.lib.gpc_create(self.kind, self.F, self.device, C.byref(h))
        L.check(self.lib, None, rc)
        self.h = h
        self.N = 0

    # -- life cycle ---------------------------------------------------------------------
    def close(self):
        if getattr(self, "h", None):
            self.lib.gpc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        L.check(self.lib, self.h, rc)

    # -- state --------------------------------------------------------------------------
    def set_hypers(self, flat, jitter):
        flat = L.as_f64(flat).ravel()
        self._ck(self.lib.gpc_set_hypers(self.h, L.dptr(flat), flat.size, float(jitter)))

    def set_data(self, X4, y, extra_noise_diag=None):
        X4 = L.as_f64(X4)
        if X4.ndim != 2 or X4.shape[1] != 4:
            raise ValueError("X4 must be (N, 4)")
        y = L.as_f64(y).ravel()
        if y.size != X4.shape[0]:
            raise ValueError("X and y disagree on N")
        e = None
        if extra_noise_diag is not None:
            e = L.as_f64(extra_noise_diag).ravel()
            if e.size != y.size:
                raise ValueError("extra_noise_diag must have N entries")
        self._ck(self.lib.gpc_set_data(self.h, L.dptr(X4), L.dptr(y), L.dptr(e), X4.shape[0]))
        self.N = X4.shape[0]

    def factor(self):
        """Returns (nlml, logdet); raises ``numpy.linalg.LinAlgError`` when not PD."""
        nlml, logdet = C.c_double(), C.c_double()
        self._ck(self.lib.gpc_factor(self.h, C.byref(nlml), C.byref(logdet)))
        self.nlml_, self.logdet_ = nlml.value, logdet.value
        return nlml.value, logdet.value

    def nlml_grad(self, n_hyp, want_diag=False):
        """Gradient of the NLML w.r.t. the flat hyper vector (and diag(Ky^-1 - alpha alpha^T))."""
        g = np.empty(int(n_hyp))
        d = np.empty(self.N) if want_diag else None
        self._ck(self.lib.gpc_nlml_grad(self.h, L.dptr(g), g.size, L.dptr(d)))
        return (g, d) if want_diag else g

    def alpha(self):
        a = np.empty(self.N)
        self._ck(self.lib.gpc_get_alpha(self.h, L.dptr(a)))
        return a

    def chol(self):
        a = np.empty((self.N, self.N))
        self._ck(self.lib.gpc_get_chol(self.h, L.dptr(a)))
        return a

    def linv(self):
        a = np.empty((self.N, self.N))
        self._ck(self.lib.gpc_get_linv(self.h, L.dptr(a)))
        return a

    def padded_n(self):
        return int(self.lib.gpc_padded_n(self.h))

    def factor_state_dev(self):
        """Device pointers (L, Linv, alpha) and n_pad -- for the NCCL broadcast of the factor."""
        pL, pX, pa, n = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_long()
        self._ck(self.lib.gpc_factor_state_dev(self.h, C.byref(pL), C.byref(pX), C.byref(pa), C.byref(n)))
        return pL.value, pX.value, pa.value, n.value

    def adopt_factor(self, logdet):
        self._ck(self.lib.gpc_adopt_factor(self.h, float(logdet)))

    # -- kernel matrix ------------------------------------------------------------------
    def kernel_matrix(self, Xa4, Xb4=None):
        Xa4 = L.as_f64(Xa4)
        Xb = None if Xb4 is None else L.as_f64(Xb4)
        nb = Xa4.shape[0] if Xb is None else Xb.shape[0]
        K = np.empty((Xa4.shape[0], nb))
        if K.size == 0:
            return K
        self._ck(self.lib.gpc_kernel_matrix(self.h, L.dptr(Xa4), Xa4.shape[0], L.dptr(Xb), nb, L.dptr(K)))
        return K

    # -- posterior ----------------------------------------------------------------------
    def predict(self, Xs4, flags, want_var=True):
        Xs4 = L.as_f64(Xs4)
        M = Xs4.shape[0]
        mean = np.empty(M)
        var = np.empty(M) if want_var else None
        if not want_var:
            flags |= L.MEAN_ONLY
        self._ck(self.lib.gpc_predict(self.h, L.dptr(Xs4), M, L.dptr(mean), L.dptr(var), flags))
        return mean, var

    def predict_dev(self, dXs4, M, dmean, dvar, flags):
        """Device-pointer variant (asynchronous on ``stream()``); pointers are ints."""
        self._ck(self.lib.gpc_predict_dev(self.h, dXs4, M, dmean, dvar, flags))

    def predict_grid_mean(self, ax, ay, az, fid=0):
        """Posterior mean on the tensor grid ``np.meshgrid(ax, ay, az, indexing="ij")`` at fidelity index ``fid``
        (squared-exponential kernels): returns an array of shape ``(len(ax), len(ay), len(az))``.  The contraction
        over the training points runs as GEMMs on the FP64 tensor cores (2 M N flop instead of M N kernel
        evaluations)."""
        ax, ay, az = (np.ascontiguousarray(np.asarray(a, dtype=np.float64).ravel()) for a in (ax, ay, az))
        out = np.empty((ax.size, ay.size, az.size))
        self._ck(self.lib.gpc_predict_grid_mean(self.h, L.dptr(ax), ax.size, L.dptr(ay), ay.size, L.dptr(az), az.size,
                                                float(fid), L.dptr(out)))
        return out

    def predict_grid_mean_dev(self, ax, ay, az, fid, dmean):
        """Device-output variant (``dmean``: int device pointer to nx * ny * nz doubles; asynchronous on ``stream()``)."""
        ax, ay, az = (np.ascontiguousarray(np.asarray(a, dtype=np.float64).ravel()) for a in (ax, ay, az))
        self._ck(self.lib.gpc_predict_grid_mean_dev(self.h, L.dptr(ax), ax.size, L.dptr(ay), ay.size, L.dptr(az), az.size,
                                                    float(fid), dmean))

    def predict_cov(self, Xs4, flags, extra_diag=None, want_mean=True):
        Xs4 = L.as_f64(Xs4)
        M = Xs4.shape[0]
        mean = np.empty(M) if want_mean else None
        cov = np.empty((M, M))
        e = None if extra_diag is None else L.as_f64(extra_diag).ravel()
        self._ck(self.lib.gpc_predict_cov(self.h, L.dptr(Xs4), M, L.dptr(mean), L.dptr(cov), L.dptr(e), flags))
        return mean, cov

    def mean_grad(self, Xs4):
        Xs4 = L.as_f64(Xs4)
        M = Xs4.shape[0]
        mean, grads = np.empty(M), np.empty((M, 3))
        self._ck(self.lib.gpc_mean_grad(self.h, L.dptr(Xs4), M, L.dptr(mean), L.dptr(grads)))
        return mean, grads

    # -- information gain ---------------------------------------------------------------
    @staticmethod
    def _ragged(cands):
        """list of (k_c, 4) arrays -> (rows (sum k, 4), offsets (C + 1,))."""
        offs = np.zeros(len(cands) + 1, dtype=np.int64)
        for i, c in enumerate(cands):
            offs[i + 1] = offs[i] + len(c)
        rows = np.zeros((max(int(offs[-1]), 1), 4))
        if offs[-1]:
            rows[:offs[-1]] = np.concatenate([np.asarray(c, float).reshape(-1, 4) for c in cands if len(c)])
        return rows, offs

    def ig_seq(self, rows4, offsets, sig_n, pred_fid=-1, flags=0, row_mask=None):
        rows4 = L.as_f64(rows4)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        Cn = offsets.size - 1
        I = np.empty(max(Cn, 1))
        best = C.c_long(-1)
        m = None if row_mask is None else np.ascontiguousarray(row_mask, dtype=np.uint8)
        self._ck(self.lib.gpc_ig_seq(self.h, L.dptr(rows4), L.lptr(offsets), Cn, float(sig_n), int(pred_fid),
                                     int(flags), L.ubptr(m), L.dptr(I), C.byref(best)))
        return I[:Cn], best.value

    def ig_selfgrid(self, rows4, offsets, pred_fid=-1, clip=True):
        rows4 = L.as_f64(rows4)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        Cn = offsets.size - 1
        I = np.empty(max(Cn, 1))
        best = C.c_long(-1)
        self._ck(self.lib.gpc_ig_selfgrid(self.h, L.dptr(rows4), L.lptr(offsets), Cn, int(pred_fid),
                                          L.CLIP_COV if clip else 0, L.dptr(I), C.byref(best)))
        return I[:Cn], best.value

    def ig_logdet(self, grid4, rows4, offsets, clip=False):
        """clip=True: both determinants are taken of covariances clipped element-wise at 1e-10
        (emukit ``predict_covariance``; ``calculatePathInfoEmuBatch``)."""
        grid4 = L.as_f64(grid4)
        rows4 = L.as_f64(rows4)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        Cn = offsets.size - 1
        I = np.empty(max(Cn, 1))
        best, prior = C.c_long(-1), C.c_double()
        self._ck(self.lib.gpc_ig_logdet_ex(self.h, L.dptr(grid4), grid4.shape[0], L.dptr(rows4), L.lptr(offsets), Cn,
                                           L.CLIP_COV if clip else 0, L.dptr(I), C.byref(prior), C.byref(best)))
        return I[:Cn], prior.value, best.value

    # -- evaluator ------------------------------------------------------------------------
    def spd_stats(self, cov, e=None):
        """(e^T inv(cov) e, ||inv(cov)||_F, logdet cov) via a device Cholesky."""
        cov = L.as_f64(cov)
        M = cov.shape[0]
        if cov.shape != (M, M):
            raise ValueError("cov must be square")
        ev = None if e is None else L.as_f64(e).ravel()
        if ev is not None and ev.size != M:
            raise ValueError("e must have M entries")
        quad, fro, ld = C.c_double(), C.c_double(), C.c_double()
        self._ck(self.lib.gpc_spd_stats(self.h, L.dptr(cov), M, L.dptr(ev), C.byref(quad), C.byref(fro), C.byref(ld)))
        return quad.value, fro.value, ld.value

    # -- measurement --------------------------------------------------------------------
    def stream(self):
        return self.lib.gpc_stream(self.h)

    def launch_count(self):
        return int(self.lib.gpc_launch_count(self.h))

    def set_mode(self, mode):
        """``_lib.MODE_INT8`` (default: tcgen05 INT8 Ozaki contraction, FP64 results), ``_lib.MODE_FP64`` (DMMA) or
        ``_lib.MODE_INT8_F32`` (the optional FP32-tolerance mode: 4 of the 6 digits, 10 of the 21 digit GEMMs)."""
        self._ck(self.lib.gpc_set_mode(self.h, int(mode)))

    def mode(self):
        return int(self.lib.gpc_get_mode(self.h))

    def set_chunk(self, m):
        self._ck(self.lib.gpc_set_chunk(self.h, int(m)))

    def enable_hot_timing(self, on=True):
        self._ck(self.lib.gpc_enable_hot_timing(self.h, int(bool(on))))

    def last_call_device_ms(self):
        """Device time (CUDA events) of the most recent information-gain call."""
        ms = C.c_double()
        self._ck(self.lib.gpc_last_call_device_ms(self.h, C.byref(ms)))
        return ms.value

    def hot_kernel_time(self, reset=False):
        ms, n, fl = C.c_double(), C.c_long(), C.c_double()
        self._ck(self.lib.gpc_hot_kernel_time(self.h, C.byref(ms), C.byref(n), C.byref(fl), int(reset)))
        return ms.value, n.value, fl.value
