import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100) GPU and the built libspyb200.so")


def nerr(got, want):
    """normwise error max|got-want| / max|want| (the parity metric of BASELINE.md section 3)."""
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape, f"shape {got.shape} != {want.shape}"
    return float(np.abs(got.astype(np.complex128) - want.astype(np.complex128)).max()
                 / max(float(np.abs(want).max()), 1e-300))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    meta = json.loads(str(z["meta"]))
    return z, meta["params"]


@pytest.fixture(scope="session")
def engine():
    from syncopy_b200.engine import get_engine
    return get_engine(0)
