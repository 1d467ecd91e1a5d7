import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def aa():
    """The product package (ctypes binding over libaa_gpu.so); builds the .so if missing."""
    pkg = importlib.import_module("audio-analyzer-rs_b200")
    if not os.path.exists(pkg.lib_path()):
        pkg.build_native()
    return pkg


@pytest.fixture(scope="session")
def O():
    """The CPU oracle (test infrastructure)."""
    from oracle import aa_oracle_py

    aa_oracle_py.build()
    return aa_oracle_py


@pytest.fixture(scope="session")
def torch_cuda():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch
