import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


# The parity gate of BASELINE.json's north_star ("within 1e-5 relative float32 tolerance"), in the scaled form
# SURVEY.md section 7 derives: |a-b| <= 1e-5 * max(|b|, 1).
RTOL = ATOL = 1e-5


def scaled_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.abs(a - b) / np.maximum(np.abs(b), 1.0)


def assert_close(a, b, what="", tol=RTOL):
    assert np.shape(a) == np.shape(b), (what, np.shape(a), np.shape(b))
    e = scaled_err(a, b)
    bad = ~(e <= tol)  # also catches NaN
    bad &= ~(np.isnan(np.asarray(a, np.float64)) & np.isnan(np.asarray(b, np.float64)))
    if bad.any():
        idx = np.argwhere(bad)[0]
        raise AssertionError(f"{what}: {bad.sum()} / {bad.size} over tol {tol:g}; worst {np.nanmax(e):.3e}; "
                             f"first at {tuple(idx)}: got {np.asarray(a)[tuple(idx)]!r} want {np.asarray(b)[tuple(idx)]!r}")


@pytest.fixture(scope="session")
def tracks():
    k = golden("kat")
    return {v: (k[f"{v}_gate_pos"], k[f"{v}_gate_yaw"], k[f"{v}_start_pos"]) for v in ("e2e", "indi")}
