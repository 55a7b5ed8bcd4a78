"""gpcore -- B200 (sm_100a) dense Gaussian-process inference core behind the reference's
``fit / predict / predict_covariance / CalcCost`` call surface.

    gpcore.nigp           drop-in for the reference ``NIGP.py``
    gpcore.GPy            the slice of GPy the reference uses (kern.RBF/Matern32, models.GPRegression,
                          likelihoods.Gaussian)
    gpcore.emukit         the slice of emukit it uses (multi_fidelity.{kernels,models,
                          convert_lists_to_array}, model_wrappers.gpy_model_wrappers)
    gpcore.infogain       batched information-gain path-cost operators (agent.CalcCost slot)
    gpcore.sharding       one-process-per-GPU sharding of test points / candidates (torch.distributed)
    gpcore.core.GPCore    the raw handle over the C ABI (include/gpcore.h)

Every numerical step runs in ``libgpcore.so`` (hand-written CUDA); there is no CPU fallback.
"""
from . import _lib
from ._lib import GpcoreError, build, load
from .core import GPCore, to_x4

__all__ = ["GPCore", "GpcoreError", "build", "load", "to_x4"]
