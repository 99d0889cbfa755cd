"""Host <-> device copy bandwidth of the GPU box, the quantity that separates bench.py's `e2e` from `value`:
pinned vs pageable, message size, both directions at once, and the effect of binding the process to the CPUs that
`nvidia-smi topo` reports as local to the GPU.  Prints one JSON line per case.

    python benchmarks/pcie_probe.py [--mb 2048]
"""
import argparse
import json
import os
import subprocess
import sys
import time

import torch


def bw(fn, nbytes, reps=3):
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return nbytes / best / 1e9


def gpu_cpu_affinity(index=0):
    try:
        out = subprocess.run(["nvidia-smi", "topo", "-C", "-i", str(index)], capture_output=True, text=True).stdout
        for tok in out.replace(",", " ").split():
            if "-" in tok and tok.replace("-", "").isdigit():
                a, b = tok.split("-")
                return set(range(int(a), int(b) + 1))
    except Exception:
        pass
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=2048)
    args = ap.parse_args()
    n = args.mb * (1 << 20) // 8
    dev = torch.empty(n, dtype=torch.float64, device="cuda")
    dev2 = torch.empty(n, dtype=torch.float64, device="cuda")
    print(json.dumps({"cpus": os.cpu_count(), "affinity_now": len(os.sched_getaffinity(0)),
                      "topo": subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout[-1500:]}))
    for label, cpus in (("default", None), ("gpu-local", gpu_cpu_affinity(0))):
        if label != "default":
            if not cpus:
                continue
            try:
                os.sched_setaffinity(0, cpus & os.sched_getaffinity(0) or os.sched_getaffinity(0))
            except Exception as e:  # noqa: BLE001
                print(json.dumps({"affinity_error": str(e)}))
                continue
        pinned = torch.empty(n, dtype=torch.float64, pin_memory=True)
        pinned.fill_(1.0)
        pinned2 = torch.empty(n, dtype=torch.float64, pin_memory=True)
        pinned2.fill_(2.0)
        pageable = torch.ones(n, dtype=torch.float64)
        s2 = torch.cuda.Stream()
        res = {"affinity": label, "mb": args.mb}
        res["h2d_pinned"] = bw(lambda: dev.copy_(pinned, non_blocking=True), 8 * n)
        res["d2h_pinned"] = bw(lambda: pinned.copy_(dev, non_blocking=True), 8 * n)
        res["h2d_pageable"] = bw(lambda: dev.copy_(pageable), 8 * n, reps=2)
        res["d2h_pageable"] = bw(lambda: pageable.copy_(dev), 8 * n, reps=2)

        def both():
            dev.copy_(pinned, non_blocking=True)
            with torch.cuda.stream(s2):
                pinned2.copy_(dev2, non_blocking=True)

        res["bidirectional_sum"] = bw(both, 16 * n)
        for chunk_mb in (16, 256):
            c = chunk_mb * (1 << 20) // 8

            def chunks():
                for o in range(0, n, c):
                    dev[o:o + c].copy_(pinned[o:o + c], non_blocking=True)

            res[f"h2d_pinned_chunks_{chunk_mb}mb"] = bw(chunks, 8 * n)
        print(json.dumps({k: (round(v, 2) if isinstance(v, float) else v) for k, v in res.items()}), flush=True)
        del pinned, pinned2, pageable


if __name__ == "__main__":
    main()
