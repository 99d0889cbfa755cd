"""sh_scipy_nk.py on the engine: the script's loop (sh_scipy_nk.py:51-61) with the residual closure and its globals replaced
by a device residual handle.  Same parameters (:15-29), same `newton_krylov(residual, Uo, verbose=1)` call shape.

    python examples/sh_scipy_nk_b200.py [N] [steps]
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np

import jfnk_b200 as jf

N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
Nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
d, k, r, g = 40.0, 0.2, 0.01, 1.0                      # sh_scipy_nk.py:15-19,28-29
U = np.random.default_rng(1234).standard_normal(N * N)  # :22 (seeded here)
F = jf.SHResidual(N=N, d=d, k=k, r=r, g=g)              # :31-39 Lap, L ; :47-49 residual
for i in range(Nsteps):
    if i == 1:
        t0 = time.perf_counter()                        # (the first step pays for CUDA start-up and the context)
    F.set_prev(U)                                       # Uo = U.copy(); UoUo; UoUoUo  (:56-58)
    U = jf.newton_krylov(F, U, verbose=(i < 2))         # U = newton_krylov(residual, Uo, verbose=1)  (:61)
print(f"{Nsteps} steps on {N}x{N}: {1e3 * (time.perf_counter() - t0) / (Nsteps - 1):.3f} ms/step (NumPy arrays in and out), "
      f"|U|_inf = {np.abs(U).max():.6f}")
U = F.steps(U, Nsteps)                                  # ... or the whole loop on the device
print(f"after {2 * Nsteps} steps: mean(U) = {U.mean():.6e}")
