"""Mesh update (solve_PMA + Euler step) with dense DCT-matrix products on a non-power-of-two grid: ms per update.
    JFNK_DMMA_GEMM=1|0 python benchmarks/dense_dct_micro.py [N] [loops]     # fp64 tensor cores | CUDA cores
"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import jfnk_b200 as jf
N = int(sys.argv[1]) if len(sys.argv) > 1 else 523
loops = int(sys.argv[2]) if len(sys.argv) > 2 else 20
xi = np.linspace(-1, 1, N); X, Y = np.meshgrid(xi, xi)
Q = torch.from_numpy((0.5 * X ** 2 + 0.5 * Y ** 2 + 0.01 * np.cos(np.pi * X) * np.cos(np.pi * Y)).reshape(-1)).cuda()
U = torch.from_numpy((-0.3 * np.exp(-4 * (X ** 2 + Y ** 2))).reshape(-1)).cuda()
P = jf.PMA2Residual(N=N)
P.relax_mesh(Q, U, 1e-6, loops=2)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); a.record()
P.relax_mesh(Q, U, 1e-6, loops=loops)
b.record(); torch.cuda.synchronize()
print(json.dumps({"N": N, "dmma": os.environ.get("JFNK_DMMA_GEMM", "1"), "ms_per_mesh_update": a.elapsed_time(b) / loops}))
