import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def have_gpu():
    import torch
    return torch.cuda.is_available()


@pytest.fixture(scope="module")
def nlp_mod():
    """The product's ctypes module with the CUDA library loaded (raises when liblpopc_b200.so is missing)."""
    from lpopc_b200 import nlp
    nlp.load_library()
    return nlp
