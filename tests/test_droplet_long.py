"""BASELINE config 2: droplet coalescence from the initdrop_coal_1_91-61 state, evolve_with_PDE's loop
(droplet.py:360-411) with the adaptive time step `scale += exp(-10 ||dU||)`: per step Newton-Krylov
(maxiter=20, f_tol=1e-7) and the mesh relaxation loop_pma(3e-9, 400), BOTH on the engine.

Three runs advance side by side, each with its own state (solution U, mesh potential Q):
  * the pure oracle (SciPy + scipy.fft), which also owns the reference time-step sequence dt_n = 1e-4 * scale;
  * engine run "lockstep": fed the oracle's dt_n each step -- the hot path (Newton-Krylov step + mesh relaxation)
    under identical inputs.  Bar: 1e-8 relative L2 (north_star) on every step until the first termination flip (the
    two sides take a different number of Newton iterations because the last iterate sits at the f_tol = 1e-7
    threshold); after a flip the two runs are on different, equally converged branches and the WORST value over the
    whole run must stay below 3e-8 (the reference itself moves by 1.8e-8 under a 1e-14 perturbation of its initial
    state, profiles/droplet_oracle_sensitivity_r1.txt).  The worst value and the number of flips are printed;
  * engine run "free": computes its own scale += exp(-10 ||dU||).  The adaptive step feeds every field difference
    back into dt (a 1e-9 field difference moves dt by ~1e-7), so this run is held to 1e-7 / 1e-6 and its deviation is
    reported.
Length: BASELINE config 2 is 100 steps, and that is what runs on the GPU (`-m gpu`, about a minute, most of it the
SciPy side).  The CPU test double of the host-logic suite runs 6 steps (it is ~100x slower than the B200);
JFNK_DROPLET_STEPS overrides both."""
import json
import os
import time

import numpy as np

import jfnk_b200 as jf
from oracle.mesh import DropletOracle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_droplet_coalescence_run(buffers):
    nsteps = int(os.environ.get("JFNK_DROPLET_STEPS", "100" if buffers.name == "cuda" else "6"))
    pmaloops = int(os.environ.get("JFNK_DROPLET_PMALOOPS", "400"))
    g = np.load(os.path.join(GOLD, "droplet_91x61.npz"))
    ref = DropletOracle()  # the reference path end to end
    ref.Q = g["state_Q"].copy()
    U_ref = g["state_U"].copy()
    scale_ref = 1.0
    runs = {}
    for name in ("lockstep", "free"):
        runs[name] = {"F": jf.DropletResidual(buffers=buffers), "U": g["state_U"].copy(), "Q": g["state_Q"].copy(),
                      "scale": 1.0, "flipped": False, "flips": 0, "worst": 0.0, "worst_q": 0.0, "t_nk": 0.0, "t_pma": 0.0, "nfev": 0, "errs": []}
    t_ref_nk = t_ref_pma = 0.0
    nfev_ref = 0
    violations = []
    for s in range(nsteps):
        dt_ref = 1e-4 * scale_ref
        # --- reference (droplet.py:371-384,411)
        t0 = time.perf_counter()
        n0 = ref.nfev
        ref_hist = []
        Unew_ref = ref.step(U_ref, dt_ref, pmaloops=0, history=ref_hist)
        t1 = time.perf_counter()
        ref.loop_pma(3e-9, pmaloops)
        t2 = time.perf_counter()
        t_ref_nk += t1 - t0
        t_ref_pma += t2 - t1
        nfev_ref += ref.nfev - n0
        scale_ref += np.exp(-10 * np.linalg.norm(Unew_ref - U_ref))
        U_ref = Unew_ref
        # --- engine runs
        for name, r in runs.items():
            F = r["F"]
            dt_n = dt_ref if name == "lockstep" else 1e-4 * r["scale"]
            t0 = time.perf_counter()
            F.set_mesh(r["Q"])
            F.set_prev(r["U"], dt_n)
            Unew = jf.newton_krylov(F, r["U"], verbose=0, maxiter=20, f_tol=1e-7)
            t1 = time.perf_counter()
            r["Q"] = F.relax_mesh(r["Q"], r["U"], 3e-9, loops=pmaloops)  # sees the OLD solution, as in the script
            t2 = time.perf_counter()
            r["t_nk"] += t1 - t0
            r["t_pma"] += t2 - t1
            r["nfev"] += F.last_history["nfev"]
            r["scale"] += np.exp(-10 * np.linalg.norm(Unew - r["U"]))
            r["U"] = Unew
            err = np.linalg.norm(Unew - U_ref) / np.linalg.norm(U_ref)
            errq = np.linalg.norm(r["Q"] - ref.Q) / np.linalg.norm(ref.Q)
            r["worst"], r["worst_q"] = max(r["worst"], err), max(r["worst_q"], errq)
            r["errs"].append(float(err))
            # Termination flips: droplet.py stops at |F|inf <= 1e-7 and the last iterates sit right at that
            # threshold (e.g. step 3: reference 1.088e-7 -> one more iteration, engine 9.35e-8 -> stops; the
            # per-iteration norms differ by the FD-JVP noise).  Both results are converged to f_tol but differ by
            # O(f_tol * ||J^-1||) ~ 1e-8.  Until the first flip the bar is 1e-8; after it the two runs are on different,
            # equally valid branches and are held to 2e-7.
            if F.last_history["nit"] != len(ref_hist[0]["iters"]):
                r["flipped"] = True
                r["flips"] += 1
            # Bars (checked after the summary is printed).  north_star's 1e-8 holds while the two runs are on the same branch
            # and the reference's own sensitivity allows it: a 1e-14 relative perturbation of the oracle's initial state moves
            # ITS field by 1.8e-8 by step 36 (profiles/droplet_oracle_sensitivity_r1.txt; the adaptive step
            # scale += exp(-10 |dU|) feeds field differences back into dt), so beyond step 50 / after a flip the bar is 3e-8.
            if name == "lockstep":
                bar = 1e-8 if (not r["flipped"] and s < 50) else 3e-8
            else:
                bar = 1e-7 if s < 30 else 1e-6
            if not err < bar:
                violations.append((name, s, float(err), bar, r["flips"]))
            if not errq < 1e-8:
                violations.append((name + ":Q", s, float(errq), 1e-8, r["flips"]))
    summary = {"steps": nsteps, "pmaloops": pmaloops, "backend": buffers.name,
               "scipy_s_per_step": {"newton_krylov": t_ref_nk / nsteps, "loop_pma": t_ref_pma / nsteps},
               "scipy_f_evals_per_step": nfev_ref / nsteps, "scipy_final_scale": scale_ref}
    for name, r in runs.items():
        summary[name] = {"worst_rel_l2_U": r["worst"], "worst_rel_l2_Q": r["worst_q"],
                         "rel_l2_U_every_10th_step": r["errs"][::10], "final_scale": r["scale"], "newton_count_flips": r["flips"],
                         "engine_s_per_step": {"newton_krylov": r["t_nk"] / nsteps, "loop_pma": r["t_pma"] / nsteps},
                         "engine_f_evals_per_step": r["nfev"] / nsteps}
    summary["violations"] = violations
    print("\nDROPLET_SUMMARY " + json.dumps(summary))
    assert not violations, violations[:5]
    assert runs["lockstep"]["worst"] < 3e-8
