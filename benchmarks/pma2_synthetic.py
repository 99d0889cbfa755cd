"""BASELINE config 3: PMA2_nk Newton-Krylov on a synthetic N x N grid (default 2048^2), 1 B200 -- the whole loop of
PMA2_nk.py:80-106 per step: metric fields of the current mesh, Crank-Nicolson term, newton_krylov(residual, U.val), and the
mesh update solve_PMA() ; Q.val += dt*Q.dt (one DCT solve per step), all on the device.

State as PMA2_nk.py:68-71 (U = 0, Q = (xi^2 + eta^2)/2); the time step is scaled k = 1e-4 (dksi/0.04)^4 so that
k beta^2 / dksi^4 equals the reference's N = 51 value (with k = 1e-4 unchanged the unpreconditioned
solve is impractical at this resolution; --unscaled-k runs it anyway).

Prints per-step records and ONE summary line with the keys of bench.py's contract:
  value / unit      steps/s with the fields resident in HBM (phases timed separately: setup, newton_krylov, mesh update)
  e2e               the same step through the public API with HOST ndarrays in and out of every call
  roofline          the dominant kernel class, algorithmic bytes / CUDA-event time vs the measured HBM peak
  cpu_baseline      the oracle (SciPy path of PMA2_nk.py) on a bounded sample grid, normalised per residual evaluation
                    and per grid point and extrapolated to this grid -- labelled as such

    python benchmarks/pma2_synthetic.py [--n 2048] [--steps 5] [--cpu-n 101]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

import jfnk_b200 as jf


def initial_state(N):
    xi = np.linspace(-1, 1, N)
    X, Y = np.meshgrid(xi, xi)
    return (0.5 * X ** 2 + 0.5 * Y ** 2).reshape(-1), np.zeros(N * N)


def cpu_sample(n, k, steps):
    """the oracle's PMA2 loop on an n x n grid: seconds and residual evaluations per step (first step = warm-up)"""
    from oracle.mesh import PMA2Oracle  # timed CPU baseline only

    o = PMA2Oracle(N=n, k=k)
    U = np.zeros(n * n)
    secs, fev = [], []
    for s in range(steps + 1):
        t0 = time.perf_counter()
        n0 = o.nfev
        U = o.step(U)  # the whole loop body of PMA2_nk.py:83-106 incl. solve_PMA and the mesh update
        secs.append(time.perf_counter() - t0)
        fev.append(o.nfev - n0)
    return float(np.mean(secs[1:])), float(np.mean(fev[1:]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=2048)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--unscaled-k", action="store_true")
    ap.add_argument("--cpu-n", type=int, default=201, help="grid of the bounded CPU-reference sample (0: skip)")
    ap.add_argument("--no-mesh-update", action="store_true", help="freeze the mesh (the round-1 measurement)")
    args = ap.parse_args()
    N = args.n
    dksi = 2.0 / (N - 1)
    k = 1e-4 if args.unscaled_k else 1e-4 * (dksi / 0.04) ** 4
    Q0, U0 = initial_state(N)
    peak = 6543.1
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = json.load(open(p)).get("hbm_gbs", peak)

    F = jf.PMA2Residual(N=N, dt=k)
    Q = torch.from_numpy(Q0).cuda()
    U = torch.from_numpy(U0).cuda()
    ph = {"setup": 0.0, "newton_krylov": 0.0, "mesh_update": 0.0}
    recs = []
    for s in range(args.warmup + args.steps):
        if s == args.warmup:
            F.profile(True)
            ph = {kk: 0.0 for kk in ph}
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        F.set_mesh(Q)
        F.set_prev(U)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        Un = jf.newton_krylov(F, U, verbose=0)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        if not args.no_mesh_update:
            dtm = float(((1 + U) ** 3).min()) * k  # compute_g() * k (PMA2_nk.py:91,446-450)
            Q = F.relax_mesh(Q, U, dtm, loops=1)   # solve_PMA() with the OLD solution ; Q.val += dt*Q.dt (:94,:103)
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        U = Un
        ph["setup"] += t1 - t0; ph["newton_krylov"] += t2 - t1; ph["mesh_update"] += t3 - t2
        h = F.last_history
        recs.append({"step": s, "seconds": round(t3 - t0, 5), "newton_its": h["nit"], "f_evals": h["nfev"],
                     "f_max": float(h["f_max"][-1]), "u_min": float(U.min())})
        print(json.dumps(recs[-1]), flush=True)
    prof = F.profile_read()
    F.profile(False)
    timed = recs[args.warmup:]
    total = sum(ph.values())
    kernels = {kk: {"launches": v["launches"], "ms": round(v["ms"], 2), "GBps": round(v["bytes"] / (v["ms"] * 1e-3) / 1e9, 1) if v["ms"] > 0 else 0,
                    "frac_of_step": round(v["ms"] * 1e-3 / total, 4)} for kk, v in prof.items()}
    dom = max(prof.items(), key=lambda kv: kv[1]["ms"])
    ach = dom[1]["bytes"] / (dom[1]["ms"] * 1e-3) / 1e9

    # ---- e2e: the same loop through the public API with HOST ndarrays (every call copies in and out) ------------------
    Qh, Uh = Q.cpu().numpy(), U.cpu().numpy()
    e2e_steps = max(2, args.steps // 2)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for s in range(e2e_steps):
        F.set_mesh(Qh)
        F.set_prev(Uh)
        Unh = jf.newton_krylov(F, Uh, verbose=0)
        if not args.no_mesh_update:
            Qh = F.relax_mesh(Qh, Uh, float(np.min((1 + Uh) ** 3)) * k, loops=1)
        Uh = Unh
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / e2e_steps

    cpu = None
    if args.cpu_n > 0:
        kc = 1e-4 * ((2.0 / (args.cpu_n - 1)) / 0.04) ** 4
        sec, fev = cpu_sample(args.cpu_n, kc, 2)
        per_eval_point = sec / fev / (args.cpu_n ** 2)
        fe = float(np.mean([o["f_evals"] for o in timed]))
        cpu = {"value": 1.0 / (per_eval_point * fe * N * N), "unit": "steps/s", "cores": 1, "kind": "port",
               "sample": f"oracle.mesh.PMA2Oracle (SciPy newton_krylov + sparse operators) on a {args.cpu_n}^2 grid, k scaled the same "
                         f"way: {sec:.3f} s/step at {fev:.0f} residual evaluations/step; normalised per evaluation and per grid "
                         f"point and extrapolated to {N}^2 at {fe:.0f} evaluations/step (EXTRAPOLATED, not run at this size)",
               "measured_seconds_per_step_on_sample": sec}
    # DRAM traffic per launch of the dominant class from the committed ncu --set full capture (2048^2 only): the marching class
    # alternates its two passes, so the figure is the mean of the two kernels' dram read + write bytes
    traffic, traffic_src = None, None
    tj = os.path.join(ROOT, "profiles", "ncu_traffic_r2.json")
    if dom[0] == "mesh_march" and N == 2048 and os.path.exists(tj):
        t = json.load(open(tj))
        if "mesh_march_lap" in t and "mesh_march_pma2_jvp" in t:
            traffic = 0.5 * (t["mesh_march_lap"]["dram_bytes"] + t["mesh_march_pma2_jvp"]["dram_bytes"])
            alg = 0.5 * (t["mesh_march_lap"]["algorithmic_bytes"] + t["mesh_march_pma2_jvp"]["algorithmic_bytes"])
            traffic_src = (f"ncu --set full, mean of the Lap pass and the PMA2 pass (profiles/ncu_traffic_r2.json): "
                           f"dram read+write = {traffic / alg:.3f} x algorithmic")
    summary = {"metric": f"PMA2_nk time-steps/s, synthetic {N}^2 fp64 (BASELINE config 3)", "value": len(timed) / total, "unit": "steps/s",
               "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(timed),
               "higher_is_better": True, "dtype": "f64", "data": "synthetic",
               "config": {"workload": f"PMA2 synthetic {N}^2, k={k:.3e}, U=0, Q=(xi^2+eta^2)/2; per step: metrics + CN term, "
                                      "newton_krylov/LGMRES, solve_PMA + explicit mesh update" + (" (mesh frozen)" if args.no_mesh_update else ""),
                          "f_evals_per_step": float(np.mean([o["f_evals"] for o in timed])),
                          "newton_its_per_step": float(np.mean([o["newton_its"] for o in timed]))},
               "phases_ms_per_step": {kk: round(1e3 * v / len(timed), 3) for kk, v in ph.items()},
               "e2e": {"value": 1.0 / e2e_s, "unit": "steps/s", "call": "set_mesh/set_prev/newton_krylov/relax_mesh with host ndarrays",
                       "h2d_bytes_per_step": 8 * N * N * 5, "d2h_bytes_per_step": 8 * N * N * 2},
               "roofline": {"bound": "hbm", "kernel": dom[0], "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                            "traffic": traffic, "traffic_source": traffic_src,
                            "note": "algorithmic bytes of the class (DESIGN.md) / CUDA-event time on the launch stream"},
               "cpu_baseline": cpu, "kernels": kernels}
    print(json.dumps(summary))


if __name__ == "__main__":
    main()
