"""droplet.py evolve_with_PDE on the engine (droplet.py:360-411) from a state file of the reference's format
(`initdrop_coal_1_91-61_100_0.01_0.01_0.1_0.15.txt`: two columns U, Q) or, without an argument, from the copy of that state
in tests/golden/.

    python examples/droplet_b200.py [state.txt] [steps]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np

import jfnk_b200 as jf

if len(sys.argv) > 1 and os.path.exists(sys.argv[1]):
    U, Q = jf.load_droplet_state(sys.argv[1])             # read_from_file (:556-576)
else:
    g = np.load(os.path.join(ROOT, "tests", "golden", "droplet_91x61.npz"))
    U, Q = g["state_U"].copy(), g["state_Q"].copy()
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
F = jf.DropletResidual(Nx=91, Ny=61)                      # droplet.py:23-53
scale, t = 1.0, 0.0
for s in range(steps):
    if s == 1:
        t0 = time.perf_counter()                          # (the first step pays for CUDA start-up and the context)
    dt_n = 1e-4 * scale                                   # :371
    F.set_mesh(Q)                                         # compute_Q_spatial_ders ; J (:373-376)
    F.set_prev(U, dt_n)                                   # U.val ; P ; F = pde_rhs (:377-381)
    Unew = jf.newton_krylov(F, U, verbose=0, maxiter=20, f_tol=1e-7)  # :383
    Q = F.relax_mesh(Q, U, 3e-9, loops=400)               # loop_pma(3e-9, 400) (:384)
    scale += np.exp(-10 * np.linalg.norm(Unew - U))       # :411
    U = Unew
    t += dt_n
print(f"{steps} steps: {1e3 * (time.perf_counter() - t0) / (steps - 1):.2f} ms/step, t = {t:.4e}, max(U) = {U.max():.6f}, scale = {scale:.4f}")
