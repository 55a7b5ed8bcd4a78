from . import gpy_model_wrappers  # noqa: F401
from .gpy_model_wrappers import GPyMultiOutputWrapper  # noqa: F401
