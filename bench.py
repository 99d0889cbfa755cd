#!/usr/bin/env python
"""Headline benchmark: JFNK implicit time steps per second of Swift-Hohenberg on a 16384 x 16384 periodic
grid (BASELINE.json configs[3]; fixed mesh spacing h = 0.625, k = 0.2, r = 0.01, g = 1), 1/2/4/8 B200
with a 1-D slab decomposition over rows.

    python bench.py --gpus 1 --steps 3 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference's SciPy path on the host cores (bounded sample)

One "step" = one pass of the hot path over the whole grid: set_prev(U); U = newton_krylov(residual, U)
(sh_scipy_nk.py:53-61).  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

H = 0.625
PARAMS = dict(k=0.2, r=0.01, g=1.0)
METRIC = "JFNK time-steps/s, Swift-Hohenberg 16384^2 fp64"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 6 and r[2 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def cpu_reference_sample(n_sample, steps, warmup):
    """The reference's SciPy path (oracle/sh.py = sh_scipy_nk.py:31-61 restated around scipy.optimize.newton_krylov)
    on the host cores, on an n_sample^2 grid with the same h, k, r, g and seeded state; returns seconds/step."""
    from oracle.sh import SHOracle, seeded_state

    o = SHOracle(N=n_sample, d=H * n_sample, **PARAMS)
    U = seeded_state(n_sample)
    hist = []
    for _ in range(warmup):
        U = o.step(U)
    t0, c0 = time.perf_counter(), time.process_time()
    for _ in range(steps):
        U = o.step(U, history=hist)
    wall = time.perf_counter() - t0
    cpu_reference_sample.effective_cores = round((time.process_time() - c0) / wall, 2)  # threads actually busy
    dt = wall / max(steps, 1)
    nfev = float(np.mean([h["nfev"] for h in hist])) if hist else 0.0
    return dt, nfev


def host_threads():
    try:
        from threadpoolctl import threadpool_info

        return max([int(i.get("num_threads", 1)) for i in threadpool_info()] + [1])
    except Exception:
        return 1


def run_reference(args, full_n):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ns = args.cpu_sample_n
    dt, nfev = cpu_reference_sample(ns, args.steps, min(args.warmup, 1))
    scale = (full_n / ns) ** 2  # cost is linear in the number of grid points at fixed h
    value = 1.0 / (dt * scale)
    sample = (f"SciPy newton_krylov/LGMRES + CSR L@u on a {ns}^2 grid (same h, k, r, g, seeded IC), {dt:.2f} s/step, "
              f"{nfev:.0f} F evals/step; steps/s of the {full_n}^2 workload extrapolated linearly in grid points "
              f"(x{scale:.0f}); the {full_n}^2 CSR operator (~56 GB) does not fit the reference's path")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / value, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"swift-hohenberg {full_n}^2 periodic, h={H}, k=0.2, r=0.01, g=1 (CPU sample {ns}^2)"},
        "cpu_baseline": {"value": value, "unit": "steps/s", "cores": host_threads(),
                         "effective_cores": getattr(cpu_reference_sample, "effective_cores", None), "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--grid", type=int, default=16384, help="grid points per side (BASELINE config: 16384)")
    ap.add_argument("--cpu-sample-n", type=int, default=1024, help="grid of the bounded CPU-reference sample")
    ap.add_argument("--gs", default="cgs-ifneeded", choices=["cgs", "cgs-ifneeded", "cgs2"])
    ap.add_argument("--gs-tau", type=float, default=0.25)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    N = args.grid

    if args.impl == "reference":
        run_reference(args, N)
        return

    import torch
    import torch.distributed as dist

    import jfnk_b200 as jf

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback; use --impl reference for the SciPy path)")
    torch.cuda.set_device(local_rank)
    comm = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        comm = jf.SlabComm()
    if args.gpus != world and rank == 0:
        print(f"bench.py: --gpus {args.gpus} but WORLD_SIZE={world}; using {world}", file=sys.stderr)

    F = jf.SHResidual(N=N, d=H * N, comm=comm, gs=args.gs, gs_tau=args.gs_tau, **PARAMS)
    row0, nrows = F.rows
    n_local = nrows * N
    # synthetic state, independent of the number of ranks (row-wise counter streams), staged in pinned memory
    host = torch.empty(n_local, dtype=torch.float64).pin_memory()
    host.numpy()[:] = jf.seeded_slab_state(N, row0, nrows, seed=1234)
    U = host.to("cuda", non_blocking=True)
    torch.cuda.synchronize()
    ctx = F.context()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    hist = []
    for _ in range(args.warmup):
        F.steps(U, 1, history=hist, inplace=True)
    barrier()

    # ---- timed region: K steps, inputs resident in HBM ------------------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    F.profile(True)
    F.profile_read()
    l0 = ctx.launches()
    thist = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    F.steps(U, args.steps, history=thist, inplace=True)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = ctx.launches() - l0
    prof = F.profile_read()
    F.profile(False)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = args.steps / (ms * 1e-3)

    # ---- end-to-end through the public API with HOST buffers: H2D of the state, step, D2H of the result -----
    e2e = None
    if not args.no_e2e:
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(args.steps):
            U.copy_(host, non_blocking=True)          # pinned host -> device
            F.steps(U, 1, inplace=True)
            host.copy_(U, non_blocking=True)          # device -> pinned host (the step's result)
            torch.cuda.current_stream().synchronize()
        f1.record()
        barrier()
        t2 = torch.tensor([f0.elapsed_time(f1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        e2e = {"value": args.steps / (float(t2.item()) * 1e-3), "unit": "steps/s",
               "h2d_bytes_per_step": int(8 * N * N), "d2h_bytes_per_step": int(8 * N * N)}

    # ---- standalone stencil SpMV bandwidth on the same grid (BASELINE metric, config 5 at this size) ---------
    spmv = {}
    x = U
    y = torch.empty_like(U)
    for name, fn in (("lap", ctx.lib.jfnk_spmv_lap), ("L", ctx.lib.jfnk_spmv_sh)):
        for _ in range(3):
            ctx.check(fn(ctx.handle, ctx.buf.ptr(x), ctx.buf.ptr(y)))
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        barrier()
        s0.record()
        for _ in range(reps):
            ctx.check(fn(ctx.handle, ctx.buf.ptr(x), ctx.buf.ptr(y)))
        s1.record()
        barrier()
        tt = torch.tensor([s0.elapsed_time(s1) / reps], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        spmv[name] = 16.0 * N * N / (float(tt.item()) * 1e-3) / 1e9  # whole-job GB/s at 16 B/point

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peaks()
    # dominant kernel class by device time inside the timed region
    main_classes = {k: v for k, v in prof.items() if v["bytes"] > 0 and not k.endswith("pass2")}
    dom = max(main_classes, key=lambda k: main_classes[k]["ms"]) if main_classes else None
    roofline = None
    kernels = {}
    for k, v in prof.items():
        gbs = v["bytes"] / (v["ms"] * 1e-3) / 1e9 if v["ms"] > 0 else 0.0
        kernels[k] = {"launches": v["launches"], "ms": round(v["ms"], 3), "GBps": round(gbs, 1),
                      "frac_of_step": round(v["ms"] / (ms if ms > 0 else 1), 4)}
    if dom:
        v = prof[dom]
        ach = v["bytes"] / (v["ms"] * 1e-3) / 1e9
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic_r1.json")
        if os.path.exists(tpath):
            rec = json.load(open(tpath)).get(dom)
            if rec:  # DRAM bytes per launch = algorithmic bytes per launch x (ncu DRAM bytes / algorithmic bytes of the captured launch)
                traffic = v["bytes"] / v["launches"] * rec["dram_bytes"] / rec["algorithmic_bytes"]
                traffic_src = f"ncu --set full on {rec['kernel']}: dram read+write = {rec['dram_bytes'] / rec['algorithmic_bytes']:.3f} x algorithmic (profiles/ncu_traffic_r1.json)"
        roofline = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                    "bytes_per_launch": v["bytes"] / v["launches"], "ms_per_launch": v["ms"] / v["launches"],
                    "note": "per-rank kernel; algorithmic bytes per launch (DESIGN.md) / CUDA-event duration on the launch stream"}

    nit = float(np.mean([h["nit"] for h in thist]))
    nfev = float(np.mean([h["nfev"] for h in thist]))
    inner = float(np.mean([h["inner_iters"] for h in thist]))
    reorth = float(np.mean([h["reorth"] for h in thist]))
    line = {
        "metric": METRIC, "value": value, "unit": "steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"swift-hohenberg {N}^2 periodic, h={H}, k=0.2, r=0.01, g=1, seeded N(0,1) state",
                   "solver": f"newton_krylov/LGMRES inner_m=30 outer_k=10, gs={args.gs} tau={args.gs_tau}",
                   "parallelism": f"row slabs x{world}",
                   "collectives": ("none (one rank)" if world == 1 else
                                   "peer memory over NVLink (direct halo stores + one-shot all-reduce kernels)" if ctx.peer_memory()
                                   else "NCCL send/recv + ncclAllReduce"),
                   "l2": "inputs exceed L2 (2.1 GB per field)" if N >= 8192 else "inputs may fit L2",
                   "newton_its_per_step": nit, "f_evals_per_step": nfev, "arnoldi_its_per_step": inner,
                   "second_gs_passes_per_step": reorth},
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "kernels": kernels,
        "spmv_GBps": {k: round(v, 1) for k, v in spmv.items()},
        "spmv_frac_of_peak": {k: round(v / (peak * world), 4) for k, v in spmv.items()},
    }
    if world == 1 and not args.no_cpu_baseline:
        ns = args.cpu_sample_n
        dt, nfev_cpu = cpu_reference_sample(ns, 1, 0)
        scale = (N / ns) ** 2
        line["cpu_baseline"] = {
            "value": 1.0 / (dt * scale), "unit": "steps/s", "cores": host_threads(),
            "effective_cores": getattr(cpu_reference_sample, "effective_cores", None), "kind": "port",
            "sample": (f"oracle/sh.py (SciPy newton_krylov + CSR L@u, the reference's path) 1 step on {ns}^2, same h/k/r/g: "
                       f"{dt:.2f} s, {nfev_cpu:.0f} F evals; extrapolated linearly in grid points x{scale:.0f} to {N}^2")}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
