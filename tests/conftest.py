import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")

CASES = ("sdss", "l32", "tiny5", "tiny12", "tiny3k", "tiny8m", "tiny16", "tiny1", "edge", "train")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_finish(session):
    """The in-tree library is git-ignored: (re)build it when it is missing or older than its sources -- but only when a
    SELECTED test needs it (the GPU tests; tests/test_capi_symbols.py and the bench contract build it themselves), so
    that the host-only tests run on a machine without nvcc.  nvcc cross-compiles sm_100a without a GPU; a fresh build takes
    ~1 minute, an up-to-date one is a no-op."""
    if any(item.get_closest_marker("gpu") is not None for item in session.items):
        from qfa_b200 import _lib
        _lib.build()


def load_case(name, tag=None):
    c = dict(np.load(os.path.join(GOLD, f"case_{name}.npz")))
    c["law"] = str(c["law"])
    c["Nb"] = int(c["Nb"])
    if tag is None:
        return c
    return c, dict(np.load(os.path.join(GOLD, f"case_{name}_{tag}.npz")))


def relerr(a, b):
    """max-norm relative error with NaN placement required to match."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    na, nb = np.isnan(a), np.isnan(b)
    assert np.array_equal(na, nb), "NaN placement differs"
    if na.all():
        return 0.0
    d = np.abs(a[~na] - b[~na]).max()
    s = np.abs(b[~nb]).max()
    return d / s if s > 0 else d


@pytest.fixture(scope="session")
def cuda_model_factory():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from qfa_b200.model import QFA

    def make(c, precision):
        Npix, Nh = c["F"].shape
        m = QFA(c["Nb"], Npix - c["Nb"], Nh, torch.device("cuda:0"), tau=c["law"],
                model_params={k: c[k] for k in ("F", "Psi", "omega", "tau0", "c0", "beta")}, precision=precision)
        m.mu = torch.tensor(c["mu"])
        return m
    return make
