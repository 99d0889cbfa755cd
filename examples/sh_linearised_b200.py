"""sh_linearised.py on the engine (sh_linearised.py:51-57): the linearly-implicit step, its direct solve replaced by LGMRES to
rtol 1e-13 on the device.

    python examples/sh_linearised_b200.py [N] [steps]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np

import jfnk_b200 as jf

N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
Nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
S = jf.SHLinearised(N=N, d=40.0, k=0.2, r=0.2, g=0.0)   # sh_linearised.py:16-24
U = 0.1 * np.random.default_rng(1234).standard_normal(N * N)
Uo = U.copy()                                           # :28
U, Uo = S.steps(U, Uo, nsteps=Nsteps)                   # :51-57, Nsteps times
print(f"{Nsteps} steps on {N}x{N}: |U|_inf = {np.abs(U).max():.6f}, matvecs of the last call = {S.last_info['matvecs']}")
