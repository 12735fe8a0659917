import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


@pytest.fixture(scope="session")
def cuda_lib():
    """The C-ABI library on a machine with a GPU; GPU tests fail loudly without it."""
    import torch
    assert torch.cuda.is_available(), "gpu-marked test running without a CUDA device"
    from alphafive_b200 import _lib
    return _lib.load()
