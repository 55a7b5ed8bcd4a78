"""ctypes binding of ``libgpcore.so`` (C ABI declared in ``include/gpcore.h``).

There is no CPU fallback: if the shared library is missing ``load()`` raises, and without a
CUDA device ``gpc_create`` returns ``GPC_ERR_CUDA`` which surfaces as ``GpcoreError``.
"""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np

PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libgpcore.so")
REPO_ROOT = os.path.dirname(PKG_DIR)

GPC_OK, GPC_ERR_NOT_PD, GPC_ERR_SHAPE, GPC_ERR_CUDA, GPC_ERR_STATE, GPC_ERR_ARG = range(6)
KIND_SF_RBF, KIND_SF_MAT32, KIND_MF_AR1_RBF, KIND_MF_AR1_MAT32, KIND_NIGP = range(5)
INCLUDE_NOISE, CLIP_DIAG, CLIP_COV, NIGP_FLOOR, MEAN_ONLY = 1, 2, 4, 8, 16
IG_FIRST_PREADDED = 1
MODE_FP64, MODE_INT8, MODE_INT8_F32, MODE_INT8_L5 = 0, 1, 2, 3

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]


class GpcoreError(RuntimeError):
    """Any non-OK status other than "not positive definite"."""


class GpcoreArgError(GpcoreError, ValueError):
    """``GPC_ERR_ARG`` (e.g. a fidelity label outside [0, F)): also a ``ValueError``, which is what emukit raises."""


def sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))]


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    hdr = os.path.join(REPO_ROOT, "include", "gpcore.h")
    return any(os.path.getmtime(s) > t for s in sources() + [hdr])


def build(force=False, verbose=False):
    """Compile ``csrc/gpc_api.cu`` for sm_100a into ``libgpcore.so`` (in-tree)."""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cmd = [nvcc] + NVCC_FLAGS + ["-o", LIB_PATH, os.path.join(CSRC, "gpc_api.cu")]
    if verbose:
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    return LIB_PATH


_dp = C.POINTER(C.c_double)
_lp = C.POINTER(C.c_long)
_ubp = C.POINTER(C.c_ubyte)
_h = C.c_void_p

# name -> (restype, argtypes); mirrors include/gpcore.h one to one (tests check the symbol list)
SIGNATURES = {
    "gpc_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(_h)]),
    "gpc_destroy": (C.c_int, [_h]),
    "gpc_last_error": (C.c_char_p, [_h]),
    "gpc_version": (C.c_int, []),
    "gpc_set_hypers": (C.c_int, [_h, _dp, C.c_int, C.c_double]),
    "gpc_set_data": (C.c_int, [_h, _dp, _dp, _dp, C.c_long]),
    "gpc_factor": (C.c_int, [_h, _dp, _dp]),
    "gpc_nlml_grad": (C.c_int, [_h, _dp, C.c_int, _dp]),
    "gpc_get_alpha": (C.c_int, [_h, _dp]),
    "gpc_get_chol": (C.c_int, [_h, _dp]),
    "gpc_get_linv": (C.c_int, [_h, _dp]),
    "gpc_padded_n": (C.c_long, [_h]),
    "gpc_factor_state_dev": (C.c_int, [_h, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), _lp]),
    "gpc_adopt_factor": (C.c_int, [_h, C.c_double]),
    "gpc_kernel_matrix": (C.c_int, [_h, _dp, C.c_long, _dp, C.c_long, _dp]),
    "gpc_predict": (C.c_int, [_h, _dp, C.c_long, _dp, _dp, C.c_uint]),
    "gpc_predict_dev": (C.c_int, [_h, C.c_void_p, C.c_long, C.c_void_p, C.c_void_p, C.c_uint]),
    "gpc_predict_grid_mean": (C.c_int, [_h, _dp, C.c_long, _dp, C.c_long, _dp, C.c_long, C.c_double, _dp]),
    "gpc_predict_grid_mean_dev": (C.c_int, [_h, _dp, C.c_long, _dp, C.c_long, _dp, C.c_long, C.c_double, C.c_void_p]),
    "gpc_predict_noisy": (C.c_int, [_h, _dp, C.c_long, _dp, C.c_long, _dp, _dp, C.c_uint]),
    "gpc_predict_cov": (C.c_int, [_h, _dp, C.c_long, _dp, _dp, _dp, C.c_uint]),
    "gpc_mean_grad": (C.c_int, [_h, _dp, C.c_long, _dp, _dp]),
    "gpc_ig_seq": (C.c_int, [_h, _dp, _lp, C.c_long, C.c_double, C.c_int, C.c_uint, _ubp, _dp, _lp]),
    "gpc_ig_selfgrid": (C.c_int, [_h, _dp, _lp, C.c_long, C.c_int, C.c_uint, _dp, _lp]),
    "gpc_ig_logdet": (C.c_int, [_h, _dp, C.c_long, _dp, _lp, C.c_long, _dp, _dp, _lp]),
    "gpc_ig_logdet_ex": (C.c_int, [_h, _dp, C.c_long, _dp, _lp, C.c_long, C.c_uint, _dp, _dp, _lp]),
    "gpc_traj_points": (C.c_int, [_h, C.c_long, _lp, _dp, _lp, _dp, C.c_double, C.c_double, C.c_int, C.c_int, C.c_double,
                                  _dp, C.c_int, _dp, _dp, _lp]),
    "gpc_spd_stats": (C.c_int, [_h, _dp, C.c_long, _dp, _dp, _dp, _dp]),
    "gpc_stream": (C.c_void_p, [_h]),
    "gpc_launch_count": (C.c_long, [_h]),
    "gpc_set_mode": (C.c_int, [_h, C.c_int]),
    "gpc_get_mode": (C.c_int, [_h]),
    "gpc_set_chunk": (C.c_int, [_h, C.c_long]),
    "gpc_hot_kernel_time": (C.c_int, [_h, _dp, _lp, _dp, C.c_int]),
    "gpc_enable_hot_timing": (C.c_int, [_h, C.c_int]),
    "gpc_last_call_device_ms": (C.c_int, [_h, _dp]),
}

_lib = None


def load():
    """Load the in-tree shared library; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GpcoreError(
            "libgpcore.so is not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
            "gpcore has no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def as_f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return a


def dptr(a):
    return None if a is None else a.ctypes.data_as(_dp)


def lptr(a):
    return None if a is None else a.ctypes.data_as(_lp)


def ubptr(a):
    return None if a is None else a.ctypes.data_as(_ubp)


def check(lib, handle, rc):
    if rc == GPC_OK:
        return
    msg = lib.gpc_last_error(handle)
    msg = msg.decode() if msg else ""
    if rc == GPC_ERR_NOT_PD:
        # the reference catches LinAlgError around its factorisations (NIGP.py:156, ...MFGP.py:392)
        raise np.linalg.LinAlgError(msg or "matrix is not positive definite")
    if rc == GPC_ERR_SHAPE:
        raise ValueError(msg)
    if rc == GPC_ERR_ARG:
        raise GpcoreArgError("gpcore status %d: %s" % (rc, msg))
    raise GpcoreError("gpcore status %d: %s" % (rc, msg))
