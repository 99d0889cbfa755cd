"""PMA2_nk.py main() on the engine (PMA2_nk.py:80-106): metrics of the mesh, Crank-Nicolson term, newton_krylov, solve_PMA and
the explicit mesh update, per step.

    python examples/pma2_nk_b200.py [N] [steps]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np

import jfnk_b200 as jf

N = int(sys.argv[1]) if len(sys.argv) > 1 else 51
Nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
k = 1e-4                                                 # PMA2_nk.py:51
F = jf.PMA2Residual(N=N, dt=k)                           # :23-40
xi = np.linspace(-1, 1, N)
X, Y = np.meshgrid(xi, xi)
Q = np.reshape(0.5 * X ** 2 + 0.5 * Y ** 2, N * N)       # :68
U = np.zeros(N * N)                                      # :71
t = 0.0
for s in range(Nsteps):
    F.set_mesh(Q)                                        # compute_Q_spatial_ders ; J (:84-87)
    F.set_prev(U)                                        # U.val ; CN_term = compute_rhs_pde() (:83,:97)
    dt = min((1 + U) ** 3) * k                           # compute_g() * k (:91) -- the residual itself uses k (:66)
    Unew = jf.newton_krylov(F, U, verbose=0)             # :100
    Q = F.relax_mesh(Q, U, dt, loops=1)                  # solve_PMA() ; Q.val += dt * Q.dt (:94,:103)
    U = Unew
    t += dt
    print(f"step {s}: t = {t:.3e}, min(U) = {U.min():.6e}, Newton its = {F.last_history['nit']}, F evals = {F.last_history['nfev']}")
