import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

entry.setup_path()
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


@pytest.fixture(scope="session")
def built_lib():
    import gpcore
    gpcore.build()
    return gpcore.load()


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def normwise(a, b, scale=None):
    """max |a - b| / max(|b|, scale): the parity measure ('normwise')."""
    a, b = np.asarray(a, float), np.asarray(b, float)
    s = np.max(np.abs(b)) if scale is None else max(np.max(np.abs(b)), scale)
    return float(np.max(np.abs(a - b)) / max(s, 1e-300))
