"""The C-ABI library: loads without a GPU, exports every symbol include/jfnk.h declares, the ctypes table covers
exactly those symbols, and compute entry points fail loudly (no CPU fallback) when no CUDA device is present."""
import ctypes as C
import os
import re
import subprocess

import pytest

import jfnk_b200 as jf
from jfnk_b200 import _capi, _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "jfnk.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(jfnk_[a-z0-9_]+)\s*\(", src)) - {"jfnk_callback"})


def test_library_exports_every_declared_symbol():
    names = declared_symbols()
    assert len(names) >= 30
    lib = C.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/jfnk.h but not exported by libjfnk.so"
    assert sorted(_capi.SIGNATURES) == names, "ctypes table and header disagree"
    assert lib.jfnk_abi_version() == _capi.ABI_VERSION


def test_library_is_sm100a_and_has_tma():
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "UBLKCP" in sass, "the stencil kernel must stage rows with TMA bulk copies"


def test_no_cpu_fallback_without_a_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    L = _lib.lib()
    assert L.jfnk_device_ok() == 0
    with pytest.raises(RuntimeError, match="no usable CUDA device"):
        jf.SHResidual(N=64).context()
    with pytest.raises(RuntimeError):
        jf.CudaBuffers()


def test_config_validation_messages():
    L = _lib.lib()
    cfg = _capi.Config()
    cfg.abi_version = 99
    assert L.jfnk_workspace_bytes(C.byref(cfg)) == 0
    h = C.c_void_p()
    assert L.jfnk_create(C.byref(cfg), None, 0, C.byref(h)) == _capi.INVALID
    assert b"abi_version" in L.jfnk_last_error()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "iterative-solvers-summer-2020_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "import oracle" not in text and "from oracle" not in text and "hostsim" not in text.replace(
                    "tests/hostsim", ""), f
