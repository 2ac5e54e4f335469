import os
import sys

import pytest

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import kgo
    return kgo.Model()


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


def _emu_lib():
    import ctypes
    import subprocess
    from kinovagrasping_b200 import _lib
    d = os.path.join(ROOT, "tests", "emu")
    subprocess.check_call(["make", "-s", "-C", d, "libkg_emu.so"])
    return _lib.bind(ctypes.CDLL(os.path.join(d, "libkg_emu.so")))


@pytest.fixture(scope="session", params=[pytest.param("emu"), pytest.param("cuda", marks=pytest.mark.gpu)])
def backend(request):
    """'emu'  = TEST-ONLY host emulation of the kernel source (checks the algorithm + host logic without a GPU);
       'cuda' = the product: libkg_b200.so on a real device, called through the C ABI."""
    if request.param == "emu":
        return _emu_lib()
    from kinovagrasping_b200 import _lib
    import torch
    assert torch.cuda.is_available(), "gpu-marked test on a box without CUDA"
    return _lib.load()


@pytest.fixture(scope="session")
def card():
    from kinovagrasping_b200.modelcard import ModelCard
    return ModelCard()
