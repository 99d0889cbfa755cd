"""The reference's own Swift-Hohenberg configuration (fixed domain d = 40, k = 0.2; sh_scipy_nk.py:15-29) at grid sizes
the unpreconditioned Krylov solve cannot afford (1234 F evaluations per step at N = 256, 5826 at
N = 512), with the Fourier preconditioner M = (I/k - L/2)^-1 given as newton_krylov's inner_M (cuFFT through torch.fft).

    python benchmarks/sh_fixed_domain_precond.py [--sizes 256,512,1024,2048] [--steps 5] [--plain-upto 256]
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


def upsampled_state(N, base=64, seed=1234):
    """the reference's initial condition (standard normal on its 64 x 64 grid, sh_scipy_nk.py:16,22) interpolated
    spectrally to N x N: the same physical problem on a finer grid (white noise at the grid scale of a fine mesh is a
    different, much rougher problem: |F| ~ 1/h^4 drives the FD-Jacobian step of SciPy's heuristics below rounding)"""
    u = np.random.default_rng(seed).standard_normal((base, base))
    if N == base:
        return u.reshape(-1)
    f = np.fft.fftshift(np.fft.fft2(u))
    pad = (N - base) // 2
    Fh = np.zeros((N, N), dtype=complex)
    Fh[pad:pad + base, pad:pad + base] = f
    return (np.real(np.fft.ifft2(np.fft.ifftshift(Fh))) * (N / base) ** 2).reshape(-1)


def run(N, steps, precond):
    F = jf.SHResidual(N=N, d=40.0)
    M = jf.SHFourierPreconditioner.for_residual(F, adaptive=True) if precond else None
    U = torch.from_numpy(upsampled_state(N)).cuda()
    nfev, nit, secs = [], [], []
    for s in range(steps + 1):  # first step is the warm-up
        F.set_prev(U)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        try:
            U = jf.newton_krylov(F, U, inner_M=M, maxiter=200)
        except jf.NoConvergence:
            # |F| ~ 1/h^4 makes SciPy's FD step omega = rdiff max(1,|x|)/max(1,|F|) fall below rounding: the
            # finite-difference Jacobian is noise and Newton stalls -- a limit of the algorithm, not of the solver
            return {"N": N, "preconditioned": bool(precond), "no_convergence_within": 200, "step": s,
                    "last_f_max": float(F.last_history["f_max"][-1])}
        torch.cuda.synchronize()
        if s:
            secs.append(time.perf_counter() - t0)
            nfev.append(F.last_history["nfev"])
            nit.append(F.last_history["nit"])
    return {"N": N, "preconditioned": bool(precond), "steps_per_s": round(len(secs) / sum(secs), 2),
            "f_evals_per_step": float(np.mean(nfev)), "newton_its_per_step": float(np.mean(nit))}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="256,512")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--plain-upto", type=int, default=256, help="also run the unpreconditioned solve up to this N")
    a = ap.parse_args()
    for N in [int(s) for s in a.sizes.split(",")]:
        print(json.dumps(run(N, a.steps, True)), flush=True)
        if N <= a.plain_upto:
            print(json.dumps(run(N, a.steps, False)), flush=True)


if __name__ == "__main__":
    main()
