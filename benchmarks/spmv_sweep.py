"""BASELINE config 5: standalone matrix-free Laplacian (5-point) and L = -Lap^2 - 2Lap + (r-1)I (13-point) SpMV
bandwidth sweep, N = 1024 ... 32768, 1/2/4/8 GPUs (row slabs; launch with torchrun for > 1 GPU).

    python benchmarks/spmv_sweep.py [--sizes 1024,2048,...] [--reps 20]

GB/s = 16 B per grid point (read x, write y) x N^2 / time (whole job, max over ranks), vs the measured HBM peak.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch
import torch.distributed as dist

import jfnk_b200 as jf


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="1024,2048,4096,8192,16384,32768")
    ap.add_argument("--reps", type=int, default=20)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    comm = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
        comm = jf.SlabComm()
    peak = 6543.1
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = json.load(open(p)).get("hbm_gbs", peak)
    rows = []
    for N in [int(s) for s in args.sizes.split(",")]:
        F = jf.SHResidual(N=N, d=0.625 * N, comm=comm, inner_m=1, outer_k=0)  # minimal Krylov workspace
        ctx = F.context()
        x = torch.randn(ctx.n, dtype=torch.float64, device="cuda")
        y = torch.empty_like(x)
        res = {"N": N, "n_gpus": world}
        for name, fn in (("lap", ctx.lib.jfnk_spmv_lap), ("L", ctx.lib.jfnk_spmv_sh)):
            for _ in range(3):
                ctx.check(fn(ctx.handle, ctx.buf.ptr(x), ctx.buf.ptr(y)))
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.reps):
                ctx.check(fn(ctx.handle, ctx.buf.ptr(x), ctx.buf.ptr(y)))
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / args.reps], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            gbs = 16.0 * N * N / (ms * 1e-3) / 1e9
            res[name] = {"ms": round(ms, 4), "GBps": round(gbs, 1), "frac_of_measured_peak": round(gbs / (peak * world), 3)}
        rows.append(res)
        if rank == 0:
            print(json.dumps(res), flush=True)
        del F, ctx, x, y
        torch.cuda.empty_cache()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
