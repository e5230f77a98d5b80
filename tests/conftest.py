"""Shared test plumbing.  `-m "not gpu"` runs in the GPU-less dev container (oracle vs golden vectors,
host logic, C-ABI symbol checks); `-m gpu` runs on a B200 and compares the CUDA path, called through
the C ABI, with the oracle.  Nothing here or in the tests reads /root/reference at run time."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "diversity-recommendations_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run through gpurun)")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    out = {}
    for key in z.files:
        a = z[key]
        out[key] = torch.from_numpy(a) if a.ndim > 0 else a.item()
    return out


@pytest.fixture(scope="session")
def native():
    """The ctypes binding; on a GPU box a missing library or a non-sm_100 device is a hard failure."""
    from divrec import _native
    _native.load_library()
    return _native


@pytest.fixture(scope="session")
def dev():
    assert torch.cuda.is_available(), "gpu-marked tests need a CUDA device"
    return torch.device("cuda:0")
