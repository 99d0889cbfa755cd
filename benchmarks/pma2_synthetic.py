"""BASELINE config 3: PMA2_nk Newton-Krylov on a synthetic N x N grid (default 2048^2), 1 B200.

State as PMA2_nk.py:68-71 (U = 0, Q = (xi^2 + eta^2)/2); the time step is scaled k = 1e-4 (dksi/0.04)^4 so that
k beta^2 / dksi^4 equals the reference's N = 51 value (with k = 1e-4 unchanged the
unpreconditioned solve is impractical at this resolution).  The mesh stays fixed during the timed steps (the
DCT mesh update is SURVEY.md section 8f "next" and runs on the host in the reference's loop).

    python benchmarks/pma2_synthetic.py [--n 2048] [--steps 5]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

import jfnk_b200 as jf


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=2048)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--unscaled-k", action="store_true")
    args = ap.parse_args()
    N = args.n
    dksi = 2.0 / (N - 1)
    k = 1e-4 if args.unscaled_k else 1e-4 * (dksi / 0.04) ** 4
    xi = np.linspace(-1, 1, N)
    X, Y = np.meshgrid(xi, xi)
    Q = torch.from_numpy((0.5 * X ** 2 + 0.5 * Y ** 2).reshape(-1)).cuda()
    F = jf.PMA2Residual(N=N, dt=k)
    F.set_mesh(Q)
    U = torch.zeros(N * N, dtype=torch.float64, device="cuda")
    F.profile(True)
    out = []
    for s in range(args.steps + 1):  # first step is the warm-up
        F.set_prev(U)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        U = jf.newton_krylov(F, U, verbose=0)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        h = F.last_history
        out.append({"step": s, "seconds": round(dt, 4), "newton_its": h["nit"], "f_evals": h["nfev"],
                    "f_max": float(h["f_max"][-1]), "u_min": float(U.min())})
        print(json.dumps(out[-1]), flush=True)
    prof = F.profile_read()
    timed = out[1:]
    summary = {"config": f"PMA2 synthetic {N}^2, k={k:.3e}", "steps_per_s": len(timed) / sum(o["seconds"] for o in timed),
               "f_evals_per_step": float(np.mean([o["f_evals"] for o in timed])),
               "kernels": {kk: {"launches": v["launches"], "ms": round(v["ms"], 2),
                                "GBps": round(v["bytes"] / (v["ms"] * 1e-3) / 1e9, 1) if v["ms"] > 0 else 0} for kk, v in prof.items()}}
    print(json.dumps(summary))


if __name__ == "__main__":
    main()
