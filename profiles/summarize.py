"""Turn the raw ncu artefacts of a gpurun call (gpurun_out/) into the small text summaries committed here.

    python profiles/summarize.py r1

  launches_<round>.csv     -> profiles/launch_shares_<round>.md   (per-kernel share of the step, cold-cache/serialised)
  prof_*_<round>.ncu-rep   -> profiles/ncu_<kernel>_<round>.md    (DRAM bytes, throughput %, occupancy, stall reasons)
"""
import csv
import glob
import io
import os
import re
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
HERE = os.path.dirname(os.path.abspath(__file__))

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
]


def short(name):
    m = re.match(r"(?:void )?(?:jfnk::)?([A-Za-z0-9_]+)(<[^>]*>)?", name)
    return (m.group(1) + (m.group(2) or "")) if m else name


def launch_shares(tag):
    path = os.path.join(OUT, f"launches_{tag}.csv")
    if not os.path.exists(path):
        return
    lines = [l for l in open(path) if l.startswith('"')]
    rows = list(csv.reader(io.StringIO("".join(lines))))
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot = defaultdict(float)
    cnt = defaultdict(int)
    for r in rows[1:]:
        v = float(r[iv].replace(",", ""))
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iu], 1e-6)
        tot[short(r[ik])] += v * scale
        cnt[short(r[ik])] += 1
    total = sum(tot.values())
    with open(os.path.join(HERE, f"launch_shares_{tag}.md"), "w") as f:
        f.write(f"# ncu launch list ({tag}): `--metrics gpu__time_duration.sum --clock-control none`\n\n")
        f.write("Per-launch times are cold-cache and serialised: compare the SHARES with bench.py's `kernels` object.\n\n")
        f.write(f"captured launches: {sum(cnt.values())}, summed device time {total:.2f} ms\n\n| kernel | launches | ms | share |\n|---|---:|---:|---:|\n")
        for k in sorted(tot, key=lambda k: -tot[k]):
            f.write(f"| `{k}` | {cnt[k]} | {tot[k]:.3f} | {100 * tot[k] / total:.1f}% |\n")
    print("wrote launch_shares_%s.md" % tag)


def ncu_report(path, tag):
    base = os.path.basename(path).replace(".ncu-rep", "")
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) < 3:
        return
    hdr, units = rows[0], rows[1]
    with open(os.path.join(HERE, f"ncu_{base}.md"), "w") as f:
        f.write(f"# ncu --set full --clock-control none: {base}\n\n")
        for r in rows[2:]:
            f.write(f"## {short(r[hdr.index('Kernel Name')])}\n\n| metric | value | unit |\n|---|---:|---|\n")
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    f.write(f"| {k} | {r[i]} | {units[i]} |\n")
            f.write("\n")
    print("wrote ncu_%s.md" % base)


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
    launch_shares(tag)
    for p in sorted(glob.glob(os.path.join(OUT, f"prof_*_{tag}.ncu-rep"))):
        ncu_report(p, tag)
