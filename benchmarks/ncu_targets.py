"""Short runs of the small-grid paths for ncu captures (one kernel class per invocation is selected with ncu's -k filter):

    ncu --set full --clock-control none --import-source on -k regex:sh_cycle_kernel --launch-skip 6 -c 1 \
        -o gpurun_out/prof_shcycle64_r2 python benchmarks/ncu_targets.py sh
    ... -k regex:mesh_cycle_kernel ...  python benchmarks/ncu_targets.py droplet
    ... -k regex:pma_relax_band_kernel ... python benchmarks/ncu_targets.py droplet
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

import jfnk_b200 as jf

what = sys.argv[1] if len(sys.argv) > 1 else "sh"
if what == "sh":
    N = 64
    F = jf.SHResidual(N=N, d=40.0)
    U = torch.from_numpy(np.random.default_rng(1234).standard_normal(N * N)).cuda()
    F.steps(U, 4, inplace=True)
else:
    g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "droplet_91x61.npz"))
    F = jf.DropletResidual()
    Q = torch.from_numpy(g["state_Q"]).cuda()
    U = torch.from_numpy(g["state_U"]).cuda()
    for s in range(2):
        F.set_mesh(Q)
        F.set_prev(U, 1e-4)
        Un = jf.newton_krylov(F, U, verbose=0, maxiter=20, f_tol=1e-7)
        Q = F.relax_mesh(Q, U, 3e-9, loops=400)
        U = Un
torch.cuda.synchronize()
print("done", what)
