"""The reference's four scripts at the sizes THEY run, engine against the oracle (the SciPy path) on the same box:

  sh_scipy_nk.py      N = 64,  newton_krylov per step                       (BASELINE config 1; also 61 x 61)
  sh_linearised.py    N = 64,  one linear solve per step (spsolve there, LGMRES rtol 1e-13 here)
  PMA2_nk.py          N = 51,  metrics + CN term + newton_krylov + solve_PMA per step
  droplet.py          91 x 61, the same + loop_pma(3e-9, 400) per step      (BASELINE config 2)

Device-resident fields on the engine side (torch CUDA tensors in and out), wall-clock per step after 3 warm-up steps,
the field compared with the oracle's at the end.  One JSON line per script.

    python benchmarks/reference_sizes.py [steps]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

import jfnk_b200 as jf


def rel(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(np.asarray(b)))


def timed(fn, steps, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / steps


def main():
    from oracle.mesh import DropletOracle, PMA2Oracle  # timed CPU baselines / checkers only
    from oracle.sh import SHLinearisedOracle, SHOracle, seeded_state

    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    out = []

    # ---- sh_scipy_nk.py ---------------------------------------------------------------------------------------------
    N = 64
    U0 = seeded_state(N)
    F = jf.SHResidual(N=N, d=40.0)
    st = {"U": torch.from_numpy(U0).cuda()}
    te = timed(lambda: F.steps(st["U"], 1, inplace=True), steps)
    o = SHOracle(N=N, d=40.0)
    t0 = time.perf_counter()
    Ur = o.run(U0, steps + 3)
    tr = (time.perf_counter() - t0) / (steps + 3)
    out.append({"script": "sh_scipy_nk.py", "grid": f"{N}x{N}", "engine_ms_per_step": 1e3 * te, "scipy_ms_per_step": 1e3 * tr,
                "speedup": tr / te, "rel_l2_field": rel(st["U"].cpu().numpy(), Ur), "steps": steps + 3})

    # ---- sh_linearised.py -------------------------------------------------------------------------------------------
    S = jf.SHLinearised(N=N)
    st = {"U": torch.from_numpy(0.1 * U0).cuda(), "Uo": torch.from_numpy(0.1 * U0).cuda()}

    def lin_step():
        st["U"], st["Uo"] = S.steps(st["U"], st["Uo"], nsteps=1)

    te = timed(lin_step, steps)
    lo = SHLinearisedOracle(N=N)
    t0 = time.perf_counter()
    Ur = lo.run(0.1 * U0, steps + 3)
    tr = (time.perf_counter() - t0) / (steps + 3)
    out.append({"script": "sh_linearised.py", "grid": f"{N}x{N}", "engine_ms_per_step": 1e3 * te, "scipy_ms_per_step": 1e3 * tr,
                "speedup": tr / te, "rel_l2_field": rel(st["U"].cpu().numpy(), Ur), "steps": steps + 3,
                "note": "reference: SuperLU spsolve per step; engine: LGMRES to rtol 1e-13 through the one-launch cycle kernel"})

    # ---- PMA2_nk.py -------------------------------------------------------------------------------------------------
    N = 51
    k = 1e-4
    xi = np.linspace(-1, 1, N)
    X, Y = np.meshgrid(xi, xi)
    P = jf.PMA2Residual(N=N)
    st = {"Q": torch.from_numpy(np.reshape(0.5 * X ** 2 + 0.5 * Y ** 2, N * N)).cuda(),
          "U": torch.zeros(N * N, dtype=torch.float64, device="cuda")}

    def pma2_step():
        P.set_mesh(st["Q"])
        P.set_prev(st["U"])
        Un = jf.newton_krylov(P, st["U"], verbose=0)
        st["Q"] = P.relax_mesh(st["Q"], st["U"], float(((1 + st["U"]) ** 3).min()) * k, loops=1)
        st["U"] = Un

    te = timed(pma2_step, steps)
    po = PMA2Oracle(N=N)
    U = np.zeros(N * N)
    t0 = time.perf_counter()
    for _ in range(steps + 3):
        U = po.step(U)
    tr = (time.perf_counter() - t0) / (steps + 3)
    out.append({"script": "PMA2_nk.py", "grid": f"{N}x{N}", "engine_ms_per_step": 1e3 * te, "scipy_ms_per_step": 1e3 * tr,
                "speedup": tr / te, "rel_l2_field": rel(st["U"].cpu().numpy(), U), "rel_l2_mesh": rel(st["Q"].cpu().numpy(), po.Q),
                "steps": steps + 3})

    # ---- droplet.py -------------------------------------------------------------------------------------------------
    g = np.load(os.path.join(ROOT, "tests", "golden", "droplet_91x61.npz"))
    D = jf.DropletResidual()
    st = {"Q": torch.from_numpy(g["state_Q"]).cuda(), "U": torch.from_numpy(g["state_U"]).cuda(), "scale": 1.0}

    def droplet_step():
        D.set_mesh(st["Q"])
        D.set_prev(st["U"], 1e-4 * st["scale"])
        Un = jf.newton_krylov(D, st["U"], verbose=0, maxiter=20, f_tol=1e-7)
        st["Q"] = D.relax_mesh(st["Q"], st["U"], 3e-9, loops=400)
        st["scale"] += float(torch.exp(-10 * torch.linalg.norm(Un - st["U"])))
        st["U"] = Un

    te = timed(droplet_step, steps)
    do = DropletOracle()
    do.Q = g["state_Q"].copy()
    U, scale = g["state_U"].copy(), 1.0
    nref = min(steps + 3, 8)  # (0.6 s per step on the host)
    t0 = time.perf_counter()
    for _ in range(nref):
        Un = do.step(U, 1e-4 * scale, dtmesh=3e-9, pmaloops=400)
        scale += np.exp(-10 * np.linalg.norm(Un - U))
        U = Un
    tr = (time.perf_counter() - t0) / nref
    out.append({"script": "droplet.py", "grid": "91x61", "engine_ms_per_step": 1e3 * te, "scipy_ms_per_step": 1e3 * tr,
                "speedup": tr / te, "steps": steps + 3, "scipy_steps_timed": nref})
    for o_ in out:
        print(json.dumps(o_), flush=True)


if __name__ == "__main__":
    main()
