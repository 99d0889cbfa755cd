"""NumPy-backed buffer provider that plugs the CPU test double (libjfnk_hostsim.so) underneath the
product's Python host layer.  TEST INFRASTRUCTURE: lets the `-m "not gpu"` suite exercise the real
host logic (csrc/engine.cpp through the real C ABI and the real Python front end) without a GPU.
"""
from __future__ import annotations

import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import jfnk_b200  # noqa: E402
from jfnk_b200 import _capi  # noqa: E402

from . import build as _build  # noqa: E402

_lib = None


def lib():
    global _lib
    if _lib is None:
        path = _build.build()
        _lib = _capi.bind(path)
        _lib.hostsim_set_comm.restype = None
        _lib.hostsim_set_comm.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    return _lib


AR_FN = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(C.c_double), C.c_int, C.c_int)
HALO_FN = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double),
                      C.POINTER(C.c_double), C.c_int)


class SimBuffers:
    name = "hostsim"

    def __init__(self):
        self.lib = lib()
        self._keep = []

    def alloc(self, n):
        return np.zeros(int(n), dtype=np.float64)

    def ptr(self, a):
        return C.c_void_p(a.ctypes.data)

    def to_device(self, x):
        return np.array(np.asarray(x, dtype=np.float64).reshape(-1), copy=True)

    def to_user(self, a, like):
        return np.array(a, copy=True).reshape(np.shape(like))

    def to_numpy(self, a):
        return np.asarray(a)

    def view_for_callback(self, ptr, n, like):
        arr = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double)), shape=(int(n),))
        return np.array(arr, copy=True).reshape(np.shape(like))

    def raw_view(self, ptr, n):
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double)), shape=(int(n),))

    def synchronize(self):
        pass

    def attach_comm(self, comm, ctx):
        """gloo-backed collectives for world_size > 1 CPU tests (ring halo exchange + allreduce)."""
        import torch
        dist = comm.dist
        rank, size = comm.rank, comm.size
        prev, nxt = (rank - 1) % size, (rank + 1) % size

        def ar(_user, buf, cnt, op):
            a = np.ctypeslib.as_array(buf, shape=(cnt,))
            t = torch.from_numpy(a.copy())
            dist.all_reduce(t, op=dist.ReduceOp.SUM if op == 0 else dist.ReduceOp.MAX, group=comm.group)
            a[:] = t.numpy()

        def halo(_user, send_first, send_last, recv_top, recv_bot, cnt):
            sf = torch.from_numpy(np.ctypeslib.as_array(send_first, shape=(cnt,)).copy())
            sl = torch.from_numpy(np.ctypeslib.as_array(send_last, shape=(cnt,)).copy())
            rt = torch.empty(cnt, dtype=torch.float64)
            rb = torch.empty(cnt, dtype=torch.float64)
            # my first rows -> previous rank's bottom halo ; my last rows -> next rank's top halo
            reqs = [dist.isend(sf, prev, group=comm.group, tag=1), dist.isend(sl, nxt, group=comm.group, tag=2),
                    dist.irecv(rt, prev, group=comm.group, tag=2), dist.irecv(rb, nxt, group=comm.group, tag=1)]
            for r in reqs:
                r.wait()
            np.ctypeslib.as_array(recv_top, shape=(cnt,))[:] = rt.numpy()
            np.ctypeslib.as_array(recv_bot, shape=(cnt,))[:] = rb.numpy()

        arc, hc = AR_FN(ar), HALO_FN(halo)
        self._keep += [arc, hc]
        self.lib.hostsim_set_comm(C.cast(arc, C.c_void_p), C.cast(hc, C.c_void_p), None)
