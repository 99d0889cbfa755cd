"""BASELINE config 1 (the reference's own CPU-runnable case): Swift-Hohenberg on the script's grid, N in {61, 64},
d = 40, k = 0.2, r = 0.01, g = 1 (sh_scipy_nk.py:15-29), seeded state, 40 implicit steps.  Reports steps/s of the
engine (device-resident loop) and the field after 40 steps; the SciPy numbers come from the oracle (tests use it
as the checker; here it is only timed next to the engine, like bench.py's cpu_baseline).

    python benchmarks/config1_small_grid.py
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

import jfnk_b200 as jf


def main():
    from oracle.sh import SHOracle, seeded_state  # timed CPU baseline / checker only

    for N in (61, 64):
        U0 = seeded_state(N)
        F = jf.SHResidual(N=N, d=40.0)
        U = torch.from_numpy(U0).cuda()
        F.steps(U.clone(), 3, inplace=True)  # warm-up
        torch.cuda.synchronize()
        hist = []
        t0 = time.perf_counter()
        Ue = F.steps(U.clone(), 40, history=hist, inplace=True)
        torch.cuda.synchronize()
        te = time.perf_counter() - t0
        o = SHOracle(N=N, d=40.0)
        href = []
        t0 = time.perf_counter()
        Ur = o.run(U0, 40, history=href)
        tr = time.perf_counter() - t0
        err = np.linalg.norm(Ue.cpu().numpy() - Ur) / np.linalg.norm(Ur)
        # where the device time goes (separate, instrumented pass: CUDA events around every launch)
        F.profile(True)
        F.steps(U.clone(), 40, inplace=True)
        torch.cuda.synchronize()
        prof = {k: {"launches": v["launches"], "us_per_launch": round(1e3 * v["ms"] / max(1, v["launches"]), 2),
                    "ms_per_step": round(v["ms"] / 40, 4)} for k, v in F.profile_read().items()}
        F.profile(False)
        print(json.dumps({"N": N, "steps": 40, "engine_steps_per_s": round(40 / te, 1), "scipy_steps_per_s": round(40 / tr, 1),
                          "rel_l2_after_40_steps": err, "f_evals_per_step": {"engine": float(np.mean([h["nfev"] for h in hist])),
                                                                              "scipy": float(np.mean([h["nfev"] for h in href]))},
                          "newton_its_equal": [h["nit"] for h in hist] == [len(h["iters"]) for h in href],
                          "launches_per_step": F.context().launches() / 83.0, "kernels": prof}), flush=True)


if __name__ == "__main__":
    main()
