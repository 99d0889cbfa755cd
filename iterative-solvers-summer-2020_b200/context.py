"""Engine context: owns the device buffers (PyTorch tensors) and the ``jfnk_ctx`` handle.

PyTorch is used for buffer ownership and streams only; every number is produced by the
sm_100a kernels behind the C ABI (``include/jfnk.h``).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi
from ._capi import Config, History, NewtonOpts


class NoConvergence(Exception):
    """Raised when the nonlinear solver fails to converge within ``maxiter`` (scipy.optimize.NoConvergence,
    scipy/optimize/_nonlin.py:34-37); ``args[0]`` is the last iterate."""


class CudaBuffers:
    """fp64 CUDA buffers owned by PyTorch."""

    name = "cuda"

    def __init__(self, device=None):
        import torch
        from . import _lib

        self.torch = torch
        self.lib = _lib.require_device()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)

    def alloc(self, n):
        return self.torch.empty(int(n), dtype=self.torch.float64, device=self.device)

    def ptr(self, a):
        return C.c_void_p(a.data_ptr())

    # Host arrays of at least this many bytes come back in page-locked memory (torch's caching host allocator: the block
    # is recycled when the array is dropped, so steady-state calls allocate nothing) -- the device->host copy then runs at
    # the PCIe rate instead of through the driver's bounce buffers, and feeding the array back in (time stepping from the
    # host: U = F.steps(U, 1)) is a direct DMA as well.
    PINNED_MIN_BYTES = 1 << 20

    def pinned_array(self, shape):
        """an uninitialised fp64 NumPy array in page-locked host memory (fast host<->device copies through the public API)"""
        t = self.torch.empty(int(np.prod(shape)), dtype=self.torch.float64, pin_memory=True)
        return t.numpy().reshape(shape)

    def to_device(self, x):
        """flat fp64 device copy of a numpy array / torch tensor (always a fresh buffer)."""
        torch = self.torch
        if isinstance(x, torch.Tensor):
            return x.detach().to(device=self.device, dtype=torch.float64, copy=True).reshape(-1).contiguous()
        # (a page-locked source -- pinned_array(), or the result of an earlier call -- is recognised by the driver: the
        # synchronous copy below is then one DMA transfer, no staging)
        return torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.float64).reshape(-1))).to(self.device)

    def to_user(self, a, like):
        """SciPy's _array_like: same container kind and shape as the user's input."""
        torch = self.torch
        if isinstance(like, torch.Tensor):
            return a.reshape(like.shape).to(device=like.device, dtype=like.dtype if like.dtype.is_floating_point else torch.float64)
        if a.numel() * 8 >= self.PINNED_MIN_BYTES:
            host = torch.empty(a.numel(), dtype=torch.float64, pin_memory=True)
            host.copy_(a.reshape(-1), non_blocking=True)
            torch.cuda.current_stream(self.device).synchronize()
            out = host.numpy()  # keeps `host` alive; dropping the array returns the block to the caching host allocator
        else:
            out = a.cpu().numpy()
        return out.reshape(np.shape(like))

    def to_numpy(self, a):
        return a.detach().cpu().numpy()

    def view_for_callback(self, ptr, n, like):
        """wrap a raw device pointer handed to a C callback; NumPy users get a host copy, as SciPy would pass."""
        torch = self.torch

        class _Raw:
            __cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f8", "data": (int(ptr), False), "version": 3}

        t = torch.as_tensor(_Raw(), device=self.device)
        if isinstance(like, torch.Tensor):
            return t.reshape(like.shape)
        return t.cpu().numpy().reshape(np.shape(like))

    def raw_view(self, ptr, n):
        """a flat fp64 tensor aliasing ``n`` doubles of device memory at ``ptr`` (no copy): what a preconditioner sees"""
        torch = self.torch

        class _Raw:
            __cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f8", "data": (int(ptr), False), "version": 3}

        return torch.as_tensor(_Raw(), device=self.device)

    def attach_comm(self, comm, ctx):
        from .slab import attach_nccl

        attach_nccl(self, comm, ctx)

    def stream_ptr(self):
        return C.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def synchronize(self):
        self.torch.cuda.synchronize(self.device)


def raise_for_status(lib, rc, x_user=None):
    """Map a C-ABI status to the exception SciPy's path raises."""
    if rc == _capi.OK:
        return
    msg = lib.jfnk_last_error().decode(errors="replace")
    if rc == _capi.NO_CONVERGENCE:
        raise NoConvergence(x_user)
    if rc in (_capi.ZERO_STEP, _capi.NONFINITE, _capi.INVALID):
        raise ValueError(msg)
    raise RuntimeError(f"libjfnk error {rc}: {msg}")


class Context:
    """One ``jfnk_ctx``: a problem on a (slab of a) grid, with its Krylov workspace."""

    def __init__(self, problem, nx, ny, *, row0=0, nrows=None, rank=0, nranks=1, inner_m=30, outer_k=10,
                 gs="cgs-ifneeded", gs_tau=0.25, kernel_variant=0, buffers=None):
        self.buf = buffers if buffers is not None else CudaBuffers()
        self.lib = self.buf.lib
        if gs not in _capi.GS_MODES:
            raise ValueError(f"gs must be one of {sorted(_capi.GS_MODES)}")
        cfg = Config()
        cfg.abi_version = _capi.ABI_VERSION
        cfg.problem = problem
        cfg.nx, cfg.ny = int(nx), int(ny)
        cfg.row0 = int(row0)
        cfg.nrows = int(ny if nrows is None else nrows)
        cfg.rank, cfg.nranks = int(rank), int(nranks)
        cfg.inner_m, cfg.outer_k = int(inner_m), int(outer_k)
        cfg.gs_mode = _capi.GS_MODES[gs]
        cfg.gs_tau = float(gs_tau)
        cfg.kernel_variant = int(kernel_variant)
        cfg.stream = self.buf.stream_ptr() if hasattr(self.buf, "stream_ptr") else None
        self.cfg = cfg
        self.n = cfg.nx * cfg.nrows
        self.n_global = cfg.nx * cfg.ny
        nbytes = self.lib.jfnk_workspace_bytes(C.byref(cfg))
        if nbytes == 0:
            # let jfnk_create produce the precise message
            nbytes = 256
        self.workspace = self.buf.alloc(nbytes // 8 + 32)
        base = self.buf.ptr(self.workspace).value
        self._ws_ptr = C.c_void_p((base + 255) // 256 * 256)
        handle = C.c_void_p()
        rc = self.lib.jfnk_create(C.byref(cfg), self._ws_ptr, C.c_size_t(nbytes), C.byref(handle))
        raise_for_status(self.lib, rc)
        self.handle = handle

    def check_stream(self, what):
        """The engine enqueues on the stream that was current when the context was created; Python code that touches the
        engine's vectors in between (callback, inner_M) runs on torch's CURRENT stream.  Refuse a mismatch instead of racing."""
        if hasattr(self.buf, "stream_ptr") and self.cfg.stream is not None:
            now = self.buf.stream_ptr().value or 0
            then = self.cfg.stream or 0
            if now != then:
                raise RuntimeError(f"{what}: torch's current CUDA stream differs from the stream this context was created on; "
                                   "create the residual object (or its context) under the stream it is used with")

    def close(self):
        if getattr(self, "handle", None):
            self.lib.jfnk_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- helpers ---------------------------------------------------------------------------------
    def check(self, rc, x_user=None):
        raise_for_status(self.lib, rc, x_user)

    def vec(self, x, name="vector"):
        a = self.buf.to_device(x)
        if a.shape[0] != self.n:
            raise ValueError(f"{name} has {a.shape[0]} elements, expected {self.n}")
        return a

    def launches(self):
        return int(self.lib.jfnk_launch_count(self.handle))

    def comm_bench(self, what, count=32, field=None, reps=200):
        """microseconds per all-reduce of ``count`` scalars (what="allreduce") or per 2-row halo exchange of the device
        vector ``field`` (what="halo"), back to back on the stream; collective over the slab ranks."""
        us = C.c_double(0.0)
        w = {"allreduce": 0, "halo": 1}[what]
        self.check(self.lib.jfnk_comm_bench(self.handle, w, int(count), self.buf.ptr(field) if field is not None else None,
                                            int(reps), C.byref(us)))
        return us.value

    def peer_memory(self):
        """True when the slab ranks talk through peer memory (direct NVLink stores + one-shot all-reduce)
        rather than NCCL send/recv + ncclAllReduce."""
        return bool(self.lib.jfnk_comm_peer_memory(self.handle))

    # -- BLAS-1 building blocks of the Arnoldi process (exposed for micro-benchmarks and unit parity) ----
    def multi_dot(self, dV, nv, stride, dw):
        """host array [V_0.w, ..., V_{nv-1}.w, w.w] for nv device vectors `stride` doubles apart in dV."""
        out = np.zeros(nv + 1)
        self.check(self.lib.jfnk_multi_dot(self.handle, int(nv), self.buf.ptr(dV), C.c_size_t(int(stride)),
                                           self.buf.ptr(dw), out.ctypes.data_as(C.POINTER(C.c_double))))
        return out

    def multi_axpy(self, dV, nv, stride, coef, dw):
        """in place w -= sum_i coef[i] V_i ; returns ||w||^2."""
        coef = np.ascontiguousarray(coef, dtype=np.float64)
        n2 = C.c_double(0.0)
        self.check(self.lib.jfnk_multi_axpy(self.handle, int(nv), self.buf.ptr(dV), C.c_size_t(int(stride)),
                                            coef.ctypes.data_as(C.POINTER(C.c_double)), self.buf.ptr(dw), C.byref(n2)))
        return n2.value

    @staticmethod
    def make_opts(f_tol=None, f_rtol=None, x_tol=None, x_rtol=None, rdiff=None, maxiter=None, iter=None,
                  line_search="armijo"):
        o = NewtonOpts()
        o.f_tol = -1.0 if f_tol is None else float(f_tol)
        o.f_rtol = -1.0 if f_rtol is None else float(f_rtol)
        o.x_tol = -1.0 if x_tol is None else float(x_tol)
        o.x_rtol = -1.0 if x_rtol is None else float(x_rtol)
        o.rdiff = -1.0 if rdiff is None else float(rdiff)
        o.maxiter = -1 if maxiter is None else int(maxiter)
        o.iter = -1 if iter is None else int(iter)
        if line_search is True:
            line_search = "armijo"
        elif line_search is False:
            line_search = None
        if line_search not in (None, "armijo"):
            if line_search == "wolfe":
                raise NotImplementedError("line_search='wolfe' is not on the reference's path (armijo / None only)")
            raise ValueError("Invalid line search")
        o.line_search = 1 if line_search == "armijo" else 0
        return o


class HistoryBuffer:
    """Host-side storage behind a ``jfnk_history``."""

    def __init__(self, capacity=256):
        self.capacity = capacity
        self.f_max = np.zeros(capacity)
        self.f_l2 = np.zeros(capacity)
        self.step = np.zeros(capacity)
        self.inner = np.zeros(capacity, dtype=np.int32)
        h = History()
        h.capacity = capacity
        h.f_max = self.f_max.ctypes.data_as(C.POINTER(C.c_double))
        h.f_l2 = self.f_l2.ctypes.data_as(C.POINTER(C.c_double))
        h.step = self.step.ctypes.data_as(C.POINTER(C.c_double))
        h.inner = self.inner.ctypes.data_as(C.POINTER(C.c_int32))
        self.c = h

    def as_dict(self):
        n = min(self.c.count, self.capacity)
        return {
            "f0_max": self.c.f0_max, "f0_l2": self.c.f0_l2, "nit": int(self.c.count), "nfev": int(self.c.nfev),
            "inner_iters": int(self.c.inner_iters), "reorth": int(self.c.reorth),
            "f_max": self.f_max[:n].copy(), "f_l2": self.f_l2[:n].copy(), "step": self.step[:n].copy(),
            "inner": self.inner[:n].copy(),
        }
