import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "timeout: per-test limit (pytest-timeout)")


@pytest.fixture(scope="session")
def oracles():
    """Build the CPU checkers (restatement always; the compiled reference where /root/reference is)."""
    from oracle import bindings
    bindings.build()
    return bindings


@pytest.fixture(scope="session")
def gpu():
    """liboimgpu.so initialised on cuda:0.  Fails loudly when the extension is missing."""
    import torch
    assert torch.cuda.is_available(), "gpu-marked test without a CUDA device"
    from oim_b200 import build, lib
    build.build()
    torch.cuda.init()
    torch.zeros(1, device="cuda:0")
    lib.init([0])
    yield lib
    lib.fini()
