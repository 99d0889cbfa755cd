"""B200-native Jacobian-free Newton-Krylov engine for the implicit time stepping of
Shiakaron/Iterative-solvers-summer-2020 (Swift-Hohenberg, PMA2 and droplet models).

Host API shaped like ``scipy.optimize.newton_krylov(F, u0, method='lgmres', f_tol=...)`` with the
reference's residual / time-step call signatures; every number is computed by hand-written sm_100a
CUDA kernels reached through the C ABI of ``include/jfnk.h`` (ctypes).  PyTorch only owns buffers.
There is no CPU fallback: importing is cheap, any compute call requires ``csrc/libjfnk.so`` and a GPU.

The directory name contains hyphens; import it through the root-level alias module ``jfnk_b200``.
"""
from .context import Context, CudaBuffers, HistoryBuffer, NoConvergence
from .nonlin import newton_krylov
from .precond import SHFourierPreconditioner
from .residuals import (DropletResidual, PMA2Residual, SHLinearised, SHResidual, load_droplet_state,
                        save_droplet_state)
from .slab import SlabComm, seeded_slab_state, slab_rows

__all__ = [
    "SHFourierPreconditioner",
    "newton_krylov", "NoConvergence", "SHResidual", "SHLinearised", "PMA2Residual", "DropletResidual",
    "Context", "CudaBuffers", "HistoryBuffer", "SlabComm", "slab_rows", "seeded_slab_state",
    "load_droplet_state", "save_droplet_state",
]
