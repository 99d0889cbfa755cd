"""1-D slab decomposition over rows (the slow axis) for multi-GPU runs: one process per GPU.

Rank p owns rows [row0, row0 + nrows) of every field (u, F, all Krylov vectors).  Inside the engine the
stencil halos travel by NCCL send/recv between ring neighbours and the Krylov dot products / norms are
combined by NCCL allreduce (csrc/cuda_ops.cu); ``torch.distributed`` is only used here to agree on the
NCCL unique id.  The reference has no parallelism of any kind; this layer is new.
"""
from __future__ import annotations

import ctypes as C

import numpy as np


def slab_rows(ny: int, rank: int, size: int):
    """Contiguous, balanced block of rows of rank ``rank``: returns (row0, nrows)."""
    r0 = (ny * rank) // size
    r1 = (ny * (rank + 1)) // size
    return r0, r1 - r0


def seeded_slab_state(nx: int, row0: int, nrows: int, seed: int = 1234) -> np.ndarray:
    """Standard-normal initial state of rows [row0, row0+nrows) whose GLOBAL field does not depend on the
    number of ranks: row r is drawn from its own stream ``default_rng([seed, r])``."""
    out = np.empty((nrows, nx), dtype=np.float64)
    for i in range(nrows):
        out[i] = np.random.default_rng([seed, row0 + i]).standard_normal(nx)
    return out.reshape(-1)


def bind_host_to_gpu(device_index=None):
    """Pin this process to the CPUs that are NUMA-local to its GPU (NVML's ideal CPU affinity), so that the page-locked
    staging buffers of the host <-> device copies are allocated on the GPU's own memory node: with one process per GPU
    on a two-socket node, slabs staged on the far socket cross the inter-socket link and halve the copy rate.
    Returns the CPU set, or None when NVML / affinity control is unavailable.  ``JFNK_NUMA_BIND=0`` disables it."""
    import os

    if os.environ.get("JFNK_NUMA_BIND", "1") == "0" or not hasattr(os, "sched_setaffinity"):
        return None
    try:
        import pynvml
        import torch

        idx = torch.cuda.current_device() if device_index is None else int(device_index)
        pynvml.nvmlInit()
        try:
            # NVML enumerates by PCI bus id; go through the UUID of the torch device to be safe with CUDA_VISIBLE_DEVICES
            uuid = str(torch.cuda.get_device_properties(idx).uuid)
            h = None
            for i in range(pynvml.nvmlDeviceGetCount()):
                hi = pynvml.nvmlDeviceGetHandleByIndex(i)
                u = pynvml.nvmlDeviceGetUUID(hi)
                u = u.decode() if isinstance(u, bytes) else u
                if uuid in u:
                    h = hi
                    break
            if h is None:
                return None
            ncpu = os.cpu_count() or 1
            words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
            cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
            allowed = os.sched_getaffinity(0)
            cpus = (cpus & allowed) or None
            if cpus and cpus != allowed:
                os.sched_setaffinity(0, cpus)
            return cpus
        finally:
            pynvml.nvmlShutdown()
    except Exception:
        return None


class SlabComm:
    """The communicator of a slab-decomposed run (wraps a ``torch.distributed`` process group)."""

    def __init__(self, group=None, bind_numa=True):
        import torch.distributed as dist

        if not dist.is_initialized():
            raise RuntimeError("SlabComm needs torch.distributed.init_process_group (one process per GPU)")
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.size = dist.get_world_size(group)
        # one process per GPU: keep its host staging memory on the GPU's NUMA node (gloo / CPU runs: nothing to bind to)
        self.cpus = bind_host_to_gpu() if (bind_numa and dist.get_backend(group) == "nccl") else None

    def slab(self, ny: int):
        return slab_rows(ny, self.rank, self.size)

    def broadcast_bytes(self, payload: bytes, nbytes: int, src: int = 0) -> bytes:
        """broadcast a small byte string through the process group (CPU objects; works for nccl and gloo)."""
        box = [payload if self.rank == src else None]
        self.dist.broadcast_object_list(box, src=src, group=self.group)
        out = box[0]
        assert len(out) == nbytes
        return out

    def attach(self, ctx):
        """give the engine context its device-side communicator"""
        ctx.buf.attach_comm(self, ctx)

    # host-side gather used by tests / examples to assemble the global field on rank 0
    def gather_rows(self, local: np.ndarray, nx: int, ny: int):
        import torch

        pieces = [None] * self.size
        self.dist.all_gather_object(pieces, np.asarray(local), group=self.group)
        return np.concatenate([np.asarray(p).reshape(-1) for p in pieces]).reshape(ny * nx)


def attach_nccl(buffers, comm: SlabComm, ctx):
    """CudaBuffers.attach_comm: rank 0 creates the NCCL unique id, everyone joins (ncclCommInitRank in C)."""
    lib = ctx.lib
    idbuf = C.create_string_buffer(128)
    if comm.rank == 0:
        ctx.check(lib.jfnk_comm_unique_id(idbuf))
    raw = comm.broadcast_bytes(bytes(idbuf.raw), 128, src=0)
    idbuf = C.create_string_buffer(raw, 128)
    ctx.check(lib.jfnk_comm_init(ctx.handle, idbuf))
