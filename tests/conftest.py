import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def cuda_lib():
    """The product library, loaded; fails loudly (no skip) when it is missing on a GPU box."""
    import torch

    assert torch.cuda.is_available(), "gpu-marked test selected but no CUDA device is visible"
    from qiskit_addon_sqd_b200 import _lib

    return _lib.load()
