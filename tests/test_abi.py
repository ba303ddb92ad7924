"""The C-ABI shared library loads and exports every symbol include/spuq_sg.h declares (no GPU,
no compute calls), and the product package never imports the oracle."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    with open(os.path.join(ROOT, "include", "spuq_sg.h")) as f:
        text = f.read()
    return sorted(set(re.findall(r"SPUQ_API[^;(]*?\b(spuq_sg_\w+)\s*\(", text)))


def test_header_symbols_exported(real_lib_path):
    names = _declared()
    assert len(names) >= 25
    lib = ctypes.CDLL(real_lib_path)
    for n in names:
        assert hasattr(lib, n), "libspuq_sg.so does not export %s" % n


def test_binding_covers_header(real_lib_path):
    from spuq_b200 import _abi
    assert sorted(_abi.SIGNATURES) == _declared()
    lib = _abi.load_only(real_lib_path)
    assert lib.spuq_sg_abi_version() == 1
    assert lib.spuq_sg_reduce_scratch_bytes() > 0
    assert lib.spuq_sg_last_error() is not None


def test_emulated_library_has_same_abi(emu):
    lib = emu.lib()
    for n in _declared():
        assert hasattr(lib, n)


def test_sm100a_cubin_present(real_lib_path):
    import shutil
    import subprocess
    if not shutil.which("cuobjdump"):
        return
    out = subprocess.run(["cuobjdump", "-lelf", real_lib_path], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "spuq_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                with open(os.path.join(dirpath, fn)) as f:
                    src = f.read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), fn
                assert "oracle." not in src.replace("oracle's", ""), fn


def test_backend_fails_loudly_without_gpu(real_lib_path):
    import pytest
    import torch
    from spuq_b200 import _abi
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    _abi.reset()
    with pytest.raises(_abi.SpuqBackendError):
        _abi.lib()
