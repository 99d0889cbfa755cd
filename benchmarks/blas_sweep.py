"""Micro-benchmark of the fused BLAS-1 kernels of the Arnoldi process through the C ABI
(jfnk_multi_dot / jfnk_multi_axpy): GB/s of algorithmic traffic versus the number of basis vectors.

    python benchmarks/blas_sweep.py [--grid 8192] [--reps 5]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

import jfnk_b200 as jf


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=8192)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--nvs", default="1,2,3,4,5,6,8,9,12,13,16,17,20,24,25,28,32,36,40")
    args = ap.parse_args()
    N = args.grid
    n = N * N
    nvs = [int(x) for x in args.nvs.split(",")]
    ctx = jf.Context(0, N, N, inner_m=2, outer_k=0)
    stride = (n + 31) // 32 * 32
    V = torch.randn(max(nvs) * stride, dtype=torch.float64, device="cuda")
    w = torch.randn(n, dtype=torch.float64, device="cuda")
    out = {}
    for nv in nvs:
        coef = np.full(nv, 1e-3)
        res = {}
        for name, fn, vecs in (("mdot", lambda: ctx.multi_dot(V, nv, stride, w), nv + 1),
                               ("maxpy_sub", lambda: ctx.multi_axpy(V, nv, stride, coef, w), nv + 2)):
            fn(); fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(args.reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.reps
            res[name] = round(vecs * 8.0 * n / (ms * 1e-3) / 1e9, 1)
        out[nv] = res
        print(nv, res, flush=True)
    print(json.dumps({"grid": N, "GBps": out, "env": {k: v for k, v in os.environ.items() if k.startswith("JFNK_")}}))


if __name__ == "__main__":
    main()
