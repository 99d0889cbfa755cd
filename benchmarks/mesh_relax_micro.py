"""Time of loop_pma(3e-9, 400) (droplet.py:384) on the 91 x 61 droplet grid for the three execution modes of the mesh
relaxation: persistent cooperative kernel (default), CUDA-graph replay of the per-stage kernels (JFNK_RELAX_FUSED=0),
eager per-stage launches (JFNK_RELAX_FUSED=0 JFNK_GRAPHS=0).

    python benchmarks/mesh_relax_micro.py
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

import jfnk_b200 as jf

g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "droplet_91x61.npz"))
F = jf.DropletResidual()
Q = torch.from_numpy(g["state_Q"]).cuda()
U = torch.from_numpy(g["state_U"]).cuda()
loops = int(sys.argv[1]) if len(sys.argv) > 1 else 400
F.relax_mesh(Q, U, 3e-9, loops=loops)  # warm-up (DCT matrices, tables)
torch.cuda.synchronize()
l0 = F.context().launches()
t0 = time.perf_counter()
reps = 5
for _ in range(reps):
    Qn = F.relax_mesh(Q, U, 3e-9, loops=loops)
torch.cuda.synchronize()
ms = (time.perf_counter() - t0) / reps * 1e3
print(json.dumps({"loops": loops, "ms_per_call": round(ms, 3), "us_per_pass": round(1e3 * ms / loops, 2),
                  "launches_per_call": (F.context().launches() - l0) / reps,
                  "fused": os.environ.get("JFNK_RELAX_FUSED", "1"), "graphs": os.environ.get("JFNK_GRAPHS", "1")}))
