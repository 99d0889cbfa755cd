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


FINGERPRINT_FILE = os.path.join(ROOT, "profiles", "sh16384_fingerprint_r2.json")
N_PROBES = 1024


def field_fingerprint(U, N, row0, nrows, world, dist):
    """P-independent description of the global field held as row slabs: sum, ||.||_2, ||.||_inf and the values at
    N_PROBES fixed (seeded) grid points.  Computed outside the timed region with torch reductions (a checker, not the
    product); the slab partials are combined over the ranks, so only the summation order depends on P."""
    import torch

    acc = torch.stack([U.sum(), (U * U).sum()])
    linf = U.abs().max().reshape(1)
    idx = np.sort(np.random.default_rng(4242).integers(0, N * N, N_PROBES))
    lo, hi = row0 * N, (row0 + nrows) * N
    mine = (idx >= lo) & (idx < hi)
    vals = torch.zeros(N_PROBES, dtype=torch.float64, device=U.device)
    if mine.any():
        vals[torch.from_numpy(np.nonzero(mine)[0]).to(U.device)] = U[torch.from_numpy(idx[mine] - lo).to(U.device)]
    if world > 1:
        dist.all_reduce(acc)
        dist.all_reduce(vals)
        dist.all_reduce(linf, op=dist.ReduceOp.MAX)
    acc = acc.cpu().numpy()
    return {"sum": float(acc[0]), "l2": float(np.sqrt(acc[1])), "linf": float(linf.item()),
            "probes": vals.cpu().numpy()}


def compare_fingerprint(fp, total_steps, nits, N):
    """against the single-GPU run of the same code committed in profiles/ (same grid, same number of steps)"""
    if not os.path.exists(FINGERPRINT_FILE):
        return None, "no committed single-GPU fingerprint (profiles/sh16384_fingerprint_r2.json)"
    ref = json.load(open(FINGERPRINT_FILE))
    ent = ref.get("runs", {}).get(f"{N}:{total_steps}")
    if ent is None:
        return None, f"no committed single-GPU entry for grid {N}, {total_steps} steps (have {sorted(ref.get('runs', {}))})"
    pr = np.asarray(ent["probes"])
    out = {"rel_l2_probes": float(np.linalg.norm(fp["probes"] - pr) / np.linalg.norm(pr)),
           "rel_sum": abs(fp["sum"] - ent["sum"]) / abs(ent["sum"]),
           "rel_l2norm": abs(fp["l2"] - ent["l2"]) / ent["l2"],
           "rel_linf": abs(fp["linf"] - ent["linf"]) / ent["linf"],
           "newton_its_equal": list(nits) == list(ent.get("newton_its", [])),
           "n1_newton_its_sum": int(sum(ent.get("newton_its", []))), "newton_its_sum": int(sum(nits))}
    out["rel_l2"] = out["rel_l2_probes"]
    return out, f"vs N=1 entry {N}:{total_steps} of profiles/sh16384_fingerprint_r2.json ({N_PROBES} seeded probe points + norms)"


def full_grid_operator_check(F, ctx, U_dev, host_state, N):
    """One L@u and one Crank-Nicolson residual F(u) at the FULL benchmark size against the oracle's roll stencil
    (oracle.sh.apply_L_roll_rows / residual_roll_rows, evaluated block of rows by block of rows on the host): exercises
    the strip / row index arithmetic of the marching kernel and 2 GB vectors where no parity test reaches."""
    import torch
    from oracle.sh import apply_L_roll_rows, residual_roll_rows

    t0 = time.perf_counter()
    y = torch.empty_like(U_dev)
    ctx.check(ctx.lib.jfnk_spmv_sh(ctx.handle, ctx.buf.ptr(U_dev), ctx.buf.ptr(y)))
    Lu = y.cpu().numpy().reshape(N, N)
    u_dev = U_dev * 0.5 + 0.25
    ctx.check(ctx.lib.jfnk_set_prev(ctx.handle, ctx.buf.ptr(U_dev)))
    ctx.check(ctx.lib.jfnk_residual(ctx.handle, ctx.buf.ptr(u_dev), ctx.buf.ptr(y)))
    Fu = y.cpu().numpy().reshape(N, N)
    u = u_dev.cpu().numpy().reshape(N, N)
    del y, u_dev
    Uo = np.asarray(host_state).reshape(N, N)
    B = 512
    eL = eF = mL = mF = 0.0
    for a in range(0, N, B):
        b = min(N, a + B)
        r1 = apply_L_roll_rows(Uo, a, b, H, PARAMS["r"])
        eL, mL = max(eL, float(np.abs(Lu[a:b] - r1).max())), max(mL, float(np.abs(r1).max()))
        r2 = residual_roll_rows(u, Uo, a, b, H, PARAMS["r"], PARAMS["g"], PARAMS["k"])
        eF, mF = max(eF, float(np.abs(Fu[a:b] - r2).max())), max(mF, float(np.abs(r2).max()))
    return {"grid": N, "L_relmax_err": eL / mL, "F_relmax_err": eF / mF, "tolerance": 1e-13,
            "ok": bool(eL / mL <= 1e-13 and eF / mF <= 1e-13), "seconds": round(time.perf_counter() - t0, 1),
            "oracle": "oracle.sh.apply_L_roll_rows / residual_roll_rows (np.roll stencil, all rows, 512-row blocks)"}


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


def workload_config(N, world):
    """The `config` object of BOTH arms (the reference arm times a bounded sample of this workload and says so in
    `cpu_baseline.sample`); run-dependent counts are reported under `solver_stats`, engine choices under `engine`."""
    return {"workload": f"swift-hohenberg {N}^2 periodic, h={H}, k=0.2, r=0.01, g=1, seeded N(0,1) state",
            "solver": "newton_krylov(method='lgmres') with SciPy's defaults: inner_m=30, outer_k=10, Armijo line search, f_tol=6e-6",
            "parallelism": f"row slabs x{world}",
            "l2": "inputs exceed L2 (2.1 GB per field)" if N >= 8192 else "inputs may fit L2"}


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
        "config": workload_config(full_n, args.gpus),
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
    ap.add_argument("--no-parity", action="store_true", help="skip the full-grid operator check against the roll oracle")
    ap.add_argument("--write-fingerprint", action="store_true",
                    help="N=1 only: record this run's field fingerprint in profiles/sh16384_fingerprint_r2.json")
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
    ctx = F.context()
    host_np = ctx.buf.pinned_array(n_local)           # page-locked NumPy array (public API)
    host_np[:] = jf.seeded_slab_state(N, row0, nrows, seed=1234)
    host = torch.from_numpy(host_np)
    U = host.to("cuda", non_blocking=True)
    torch.cuda.synchronize()

    parity = {}
    if world == 1 and not args.no_parity:
        parity["full_grid_operators"] = full_grid_operator_check(F, ctx, U, host_np, N)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    hist = []
    for _ in range(args.warmup):
        F.steps(U, 1, history=hist, inplace=True)
    barrier()
    # the e2e leg below re-runs the SAME K steps from this state through the host-array API (the work per step falls as the
    # pattern forms: 6 Newton iterations in step 1, 3 by step 20 -- a leg started later would not be comparable)
    U_after_warmup = None if args.no_e2e else U.clone()

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

    # ---- parity evidence, outside the timed region: the field after warmup + timed steps ------------------------
    total_steps = args.warmup + args.steps
    nits_all = [int(h["nit"]) for h in hist + thist]
    fp = field_fingerprint(U, N, row0, nrows, world, dist if world > 1 else None)
    cmp_, cmp_note = compare_fingerprint(fp, total_steps, nits_all, N)
    parity["field_after_steps"] = {"steps_taken": total_steps, "sum": fp["sum"], "l2": fp["l2"], "linf": fp["linf"],
                                   "probe_l2": float(np.linalg.norm(fp["probes"])), "newton_its": nits_all}
    parity["vs_n1"] = cmp_
    parity["vs_n1_note"] = cmp_note
    if cmp_ is not None:
        parity["rel_l2"] = cmp_["rel_l2"]
    if args.write_fingerprint and world == 1 and rank == 0:
        db = json.load(open(FINGERPRINT_FILE)) if os.path.exists(FINGERPRINT_FILE) else {"runs": {}}
        db["what"] = ("bench.py field fingerprints of single-GPU runs: sum, l2, linf, values at 1024 seeded probe points "
                      "(default_rng(4242)), Newton iterations per step; key = grid:total_steps (warmup + timed)")
        db["runs"][f"{N}:{total_steps}"] = {"sum": fp["sum"], "l2": fp["l2"], "linf": fp["linf"],
                                            "probes": [float(v) for v in fp["probes"]], "newton_its": nits_all}
        with open(FINGERPRINT_FILE, "w") as f:
            json.dump(db, f)

    # ---- end-to-end through the PUBLIC API with HOST buffers: U = F.steps(U_host_ndarray, 1) --------------------------
    # (ndarray in -> host->device copy, set_prev + Newton-Krylov, device->host copy -> ndarray out; the call a user of the
    #  reference's loop `U = newton_krylov(residual, Uo)` makes per time step)
    e2e = None
    if not args.no_e2e:
        Uh = host_np
        host.copy_(U_after_warmup, non_blocking=True)
        torch.cuda.synchronize()
        del U_after_warmup
        # the two copies alone, to report their bandwidth
        c0, c1, c2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        c0.record(); U.copy_(host, non_blocking=True); c1.record(); host.copy_(U, non_blocking=True); c2.record()
        torch.cuda.synchronize()
        h2d_ms, d2h_ms = c0.elapsed_time(c1), c1.elapsed_time(c2)
        # warm torch's caching host allocator with the TWO page-locked result blocks the loop alternates between (a 2 GB
        # cudaHostAlloc inside the timed region costs ~0.5 s); results dropped: the timed leg starts from the same state
        _w1 = F.steps(Uh, 1)
        _w2 = F.steps(_w1, 1)
        del _w1, _w2
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            Uh = F.steps(Uh, 1)
        torch.cuda.synchronize()
        dt_e2e = time.perf_counter() - t0
        barrier()
        t2 = torch.tensor([dt_e2e], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        nb_ = int(8 * n_local)
        e2e = {"value": args.steps / float(t2.item()), "unit": "steps/s",
               "h2d_bytes_per_step": int(8 * N * N), "d2h_bytes_per_step": int(8 * N * N),
               "call": "U = SHResidual.steps(U_ndarray, 1)  (page-locked NumPy arrays in and out)",
               "h2d_ms": round(h2d_ms, 2), "d2h_ms": round(d2h_ms, 2),
               "h2d_GBps": round(nb_ / (h2d_ms * 1e-3) / 1e9, 1), "d2h_GBps": round(nb_ / (d2h_ms * 1e-3) / 1e9, 1),
               "host_numa_bound_cpus": (len(comm.cpus) if (comm is not None and comm.cpus) else None)}
        U = torch.from_numpy(Uh).to("cuda")

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
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic_r2.json")
        if os.path.exists(tpath):
            rec = json.load(open(tpath)).get(dom)
            if rec:  # DRAM bytes per launch = algorithmic bytes per launch x (ncu DRAM bytes / algorithmic bytes of the captured launch)
                traffic = v["bytes"] / v["launches"] * rec["dram_bytes"] / rec["algorithmic_bytes"]
                traffic_src = f"ncu --set full on {rec['kernel']}: dram read+write = {rec['dram_bytes'] / rec['algorithmic_bytes']:.3f} x algorithmic (profiles/ncu_traffic_r2.json)"
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
        "config": workload_config(N, world),
        "engine": {"gram_schmidt": f"{args.gs} tau={args.gs_tau}",
                   "collectives": ("none (one rank)" if world == 1 else
                                   "peer memory over NVLink (direct halo stores + all-reduce fused into the producing kernels)"
                                   if ctx.peer_memory() else "NCCL send/recv + ncclAllReduce")},
        "solver_stats": {"newton_its_per_step": nit, "f_evals_per_step": nfev, "arnoldi_its_per_step": inner,
                         "second_gs_passes_per_step": reorth},
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "parity": parity,
        "kernels": kernels,
        "spmv_GBps": {k: round(v, 1) for k, v in spmv.items()},
        "spmv_frac_of_peak": {k: round(v / (peak * world), 4) for k, v in spmv.items()},
    }
    if world == 1 and not args.no_cpu_baseline:
        ns = args.cpu_sample_n
        dt, nfev_cpu = cpu_reference_sample(ns, 2, 1)  # one warm-up step, two timed (the GPU arm is timed after warm-up too)
        scale = (N / ns) ** 2
        line["cpu_baseline"] = {
            "value": 1.0 / (dt * scale), "unit": "steps/s", "cores": host_threads(),
            "effective_cores": getattr(cpu_reference_sample, "effective_cores", None), "kind": "port",
            "sample": (f"oracle/sh.py (SciPy newton_krylov + CSR L@u, the reference's path) 2 steps after 1 warm-up step on {ns}^2, same h/k/r/g: "
                       f"{dt:.2f} s/step, {nfev_cpu:.0f} F evals/step; extrapolated linearly in grid points x{scale:.0f} to {N}^2")}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
