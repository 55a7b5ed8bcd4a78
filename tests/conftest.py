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


def measured(name, value, tol):
    """Record a measured parity figure next to the tolerance it is asserted against (printed with -s and appended to
    gpurun_out/parity_r02.jsonl), then return it: ``assert measured("ig_sf_ld", normwise(a, b), 1e-9) < 1e-9``."""
    import json
    path = os.environ.get("GPC_PARITY_LOG") or os.path.join(ROOT, "gpurun_out", "parity_r02.jsonl")
    try:
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, "a") as f:
            f.write(json.dumps({"case": name, "measured": float(value), "tolerance": float(tol)}) + "\n")
    except OSError:
        pass
    print("measured %-46s %.3e (tolerance %.1e)" % (name, value, tol))
    return value
