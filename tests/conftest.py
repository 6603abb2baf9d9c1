import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run on the GPU box with -m gpu)")


@pytest.fixture(scope="session")
def cuda():
    """torch + a CUDA device, or a loud failure (GPU tests never fall back to CPU)."""
    import torch

    if not torch.cuda.is_available():
        pytest.fail("this test is marked gpu and needs a CUDA device; run with -m 'not gpu' on CPU boxes")
    import dtfft_b200

    assert dtfft_b200.lib().dtfftb_device_available() == 1
    torch.cuda.set_device(0)
    return torch
