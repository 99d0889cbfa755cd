"""Micro-benchmark of the fused PMA2 FD-JVP (two passes of the marching Laplace kernel, mesh_march.cuh) on an N x N grid:
device time per launch from the engine's CUDA-event profile and algorithmic GB/s (7 + 11 fields of 8 B per point).

    python benchmarks/pma2_jvp_micro.py [N]            # JFNK_MARCH_DEBUG=1|2 idles the frame / interior CTAs
"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import jfnk_b200 as jf
N = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
xi = np.linspace(-1, 1, N); X, Y = np.meshgrid(xi, xi)
Q = torch.from_numpy((0.5 * X ** 2 + 0.5 * Y ** 2).reshape(-1)).cuda()
dk = 2.0 / (N - 1)
F = jf.PMA2Residual(N=N, dt=1e-4 * (dk / 0.04) ** 4)
F.set_mesh(Q)
U = torch.zeros(N * N, dtype=torch.float64, device="cuda")
F.set_prev(U)
F.linearize(U)
v = torch.from_numpy((np.sin(2 * X) * np.cos(Y)).reshape(-1)).cuda()
for _ in range(3): F.jvp(v)
F.profile(True)
for _ in range(20): F.jvp(v)
p = F.profile_read()
m = p.get("mesh_march", p.get("mesh"))
print(os.environ.get("JFNK_MARCH_DEBUG", "0"), "mesh launches", m["launches"], "us/launch", 1e3 * m["ms"] / m["launches"], "GB/s", m["bytes"] / m["ms"] / 1e6)
