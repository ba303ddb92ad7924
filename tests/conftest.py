import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

EMU_LIB = os.path.join(ROOT, "tests", "_emu", "libspuq_sg_emu.so")
REAL_LIB = os.path.join(ROOT, "spuq_b200", "libspuq_sg.so")
CSRC = os.path.join(ROOT, "spuq_b200", "csrc")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _newest_source():
    return max(os.path.getmtime(os.path.join(CSRC, f)) for f in os.listdir(CSRC))


def _ensure(path, target):
    if not os.path.isfile(path) or os.path.getmtime(path) < _newest_source():
        subprocess.run(["make", "-C", CSRC] + ([target] if target else []), check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
    return path


@pytest.fixture(scope="session")
def real_lib_path():
    """libspuq_sg.so (cross-compiled for sm_100a; loadable without a GPU)."""
    return _ensure(REAL_LIB, None)


@pytest.fixture()
def emu():
    """Route spuq_b200 to the host-emulated build of the same CUDA sources (CPU tests only)."""
    from spuq_b200 import _abi
    _ensure(EMU_LIB, "emu")
    _abi.use_library(EMU_LIB, "cpu")
    yield _abi
    _abi.reset()


@pytest.fixture()
def gpu(real_lib_path):
    """The real library on cuda:0; fails loudly if either is missing."""
    import torch
    from spuq_b200 import _abi
    assert torch.cuda.is_available(), "gpu-marked test run without a CUDA device"
    _abi.reset()
    _abi.lib()
    assert _abi.library_path() == REAL_LIB and _abi.device().type == "cuda"
    yield _abi
    torch.cuda.synchronize()
