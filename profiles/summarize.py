#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into profiles/ (tracked).

  python profiles/summarize.py <tag> <launches.csv> <prof.ncu-rep>

Writes profiles/<tag>_launches.csv (copy), profiles/<tag>_launches.txt (per-kernel totals and
shares) and profiles/<tag>_full.txt (the metrics the roofline discussion uses).
"""
import csv
import io
import os
import shutil
import subprocess
import sys
from collections import defaultdict

HERE = os.path.dirname(os.path.abspath(__file__))
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second"]


def launches(tag, path):
    shutil.copy(path, os.path.join(HERE, f"{tag}_launches.csv"))
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot = defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        v = v / 1e6 if r[ui] == "ns" else (v / 1e3 if r[ui] == "us" else v)
        tot[r[ki]][0] += 1
        tot[r[ki]][1] += v
    total = sum(t for _, t in tot.values())
    with open(os.path.join(HERE, f"{tag}_launches.txt"), "w") as fh:
        fh.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised): shares, not absolutes\n")
        fh.write(f"# total {total:.3f} ms over {sum(n for n, _ in tot.values())} launches\n")
        for k, (n, t) in sorted(tot.items(), key=lambda x: -x[1][1]):
            fh.write(f"{t:12.3f} ms  {100 * t / total:6.2f}%  x{n:<4d} avg {t / n:10.3f} ms  {k}\n")


def full(tag, rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    with open(os.path.join(HERE, f"{tag}_full.txt"), "w") as fh:
        fh.write(f"# ncu --set full --clock-control none --import-source on; source: {os.path.basename(rep)}\n")
        for vals in rows[2:]:
            name = vals[hdr.index("Kernel Name")]
            fh.write(f"\n== {name}\n")
            stalls = []
            for h, u, v in zip(hdr, units, vals):
                if h in KEYS:
                    fh.write(f"{h:75s} {v:>22s} {u}\n")
                if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
                    try:
                        stalls.append((float(v), h))
                    except ValueError:
                        pass
            fh.write("-- warp stall reasons (warps per issue-active cycle), top 8\n")
            for v, h in sorted(stalls, reverse=True)[:8]:
                fh.write(f"{v:8.3f} {h}\n")


if __name__ == "__main__":
    tag = sys.argv[1]
    launches(tag, sys.argv[2])
    if len(sys.argv) > 3:
        full(tag, sys.argv[3])
