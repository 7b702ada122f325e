import importlib
import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
PKG_NAME = "correlated-photon-mapping-for-interactive-global-illumination-of-time-varying-volumetric-data_b200"
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def cpm():
    """the product package (ctypes binding of libcpm_b200.so)"""
    return importlib.import_module(PKG_NAME)


@pytest.fixture(scope="session")
def synth():
    return importlib.import_module(PKG_NAME + ".synth")


@pytest.fixture(scope="session")
def orc():
    """the CPU oracle (test infrastructure)"""
    from oracle import orc as o
    o.lib()
    return o


@pytest.fixture(scope="session")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("gpu test selected but no CUDA device is visible (there is no CPU fallback)")
    return torch


@pytest.fixture(scope="session")
def ctx(cpm, torch_cuda):
    """One context for the session, on a stream that is also torch's current stream: tensor initialisation
    (torch.zeros, .cuda() uploads) and the library's kernels are then ordered without explicit syncs."""
    torch = torch_cuda
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    c = cpm.Context(0, stream.cuda_stream)
    yield c
    c.close()


GOLDEN = ROOT / "tests" / "golden"
