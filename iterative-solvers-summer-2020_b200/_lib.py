"""Loader of the CUDA engine ``csrc/libjfnk.so``.  There is no fallback: if the library is
missing or cannot be loaded the import of any compute entry point raises."""
from __future__ import annotations

import os

from . import _capi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libjfnk.so")
_lib = None


def lib():
    """The bound ``libjfnk.so`` (built in-tree by ``__graft_entry__.build()`` / ``csrc/build.py``)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: the sm_100a CUDA engine has not been built "
                "(run `python __graft_entry__.py` or `python iterative-solvers-summer-2020_b200/csrc/build.py`). "
                "There is no CPU fallback.")
        _lib = _capi.bind(LIB_PATH)
    return _lib


def require_device():
    """Raise unless a CUDA device is visible and the sm_100a kernel image loads."""
    L = lib()
    if not L.jfnk_device_ok():
        raise RuntimeError("libjfnk: no usable CUDA device: " + L.jfnk_last_error().decode())
    return L
