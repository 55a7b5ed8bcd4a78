"""The slice of the GPy API the reference touches, served by the CUDA core
(``import gpcore.GPy as GPy`` in place of ``import GPy``; GPTrainers.py:9)."""
from . import kern, likelihoods, models  # noqa: F401
