"""The slice of the emukit API the reference touches (GPTrainers.py:10-13)."""
from . import model_wrappers, multi_fidelity  # noqa: F401
