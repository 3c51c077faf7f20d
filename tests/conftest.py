import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def seqj():
    import sequencerj_b200 as s
    return s


@pytest.fixture(scope="session")
def oracle():
    from oracle import seq_oracle
    return seq_oracle


@pytest.fixture(scope="session")
def handle(seqj):
    """One library handle per session; creation fails loudly without an sm_100 device."""
    return seqj.default_handle()


def rel_close(a, b, rtol=1e-9, atol=0.0):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    fin = np.isfinite(b)
    assert np.array_equal(np.isfinite(a), fin)
    err = np.abs(a[fin] - b[fin])
    tol = rtol * np.abs(b[fin]) + atol
    bad = err > tol
    if bad.any():
        k = np.argmax(err - tol)
        raise AssertionError(f"{bad.sum()} entries differ; worst |d|={err[k]:.3e} ref={b[fin][k]:.6e} "
                             f"rel={err[k] / abs(b[fin][k]):.3e}")
