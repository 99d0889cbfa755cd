"""Where a droplet time step (BASELINE config 2: 91 x 61 grid, droplet.py:370-411) spends its time on the device:
wall-clock of the three host calls of a step (set_mesh + set_prev, newton_krylov, relax_mesh) with device-resident
fields, and the per-kernel-class CUDA-event times of the Newton-Krylov part.

    python benchmarks/droplet_step_micro.py [steps]
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
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30


def run(profile):
    F = jf.DropletResidual()
    Q = torch.from_numpy(g["state_Q"]).cuda()
    U = torch.from_numpy(g["state_U"]).cuda()
    scale = 1.0
    t = {"setup": 0.0, "newton_krylov": 0.0, "relax_mesh": 0.0}
    nfev = nit = 0
    for s in range(steps + 3):
        if s == 3:  # warm-up done
            torch.cuda.synchronize()
            t = {k: 0.0 for k in t}
            nfev = nit = 0
            if profile:
                F.profile(True)
        dt = 1e-4 * scale
        t0 = time.perf_counter()
        F.set_mesh(Q)
        F.set_prev(U, dt)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        Un = jf.newton_krylov(F, U, verbose=0, maxiter=20, f_tol=1e-7)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        Q = F.relax_mesh(Q, U, 3e-9, loops=400)
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        t["setup"] += t1 - t0; t["newton_krylov"] += t2 - t1; t["relax_mesh"] += t3 - t2
        nfev += F.last_history["nfev"]; nit += F.last_history["nit"]
        scale += float(torch.exp(-10 * torch.linalg.norm(Un - U)))
        U = Un
    out = {"steps": steps, "ms_per_step": {k: round(1e3 * v / steps, 3) for k, v in t.items()},
           "f_evals_per_step": nfev / steps, "newton_its_per_step": nit / steps}
    if profile:
        out["kernels"] = {k: {"launches_per_step": v["launches"] / steps, "us_per_launch": round(1e3 * v["ms"] / max(1, v["launches"]), 2),
                              "ms_per_step": round(v["ms"] / steps, 4)} for k, v in F.profile_read().items()}
    return out


print(json.dumps({"mode": "timed", **run(False)}), flush=True)
print(json.dumps({"mode": "profiled (CUDA events around every launch)", **run(True)}), flush=True)
