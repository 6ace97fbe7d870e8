#!/usr/bin/env python
"""Development tool: warp-stall samples of the main kernel by CUDA source line of ttvb_kernels.cuh
(inlined arithmetic of ttvb_math.cuh is charged to the line that calls it).

    ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass > cs.csv
    python scripts/profile_lines.py cs.csv [top] [git-rev of the profiled source]
"""
import csv
import os
import subprocess
import sys

allrows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
# the report lists every source file in turn (an instruction appears under each line of its inline
# stack, so every file's section sums to the whole kernel): keep the section of ttvb_kernels.cuh
rows, on = [], False
for r in allrows:
    if r and r[0] == "File Path":
        on = r[1].endswith("ttvb_kernels.cuh")
        continue
    if on:
        rows.append(r)
hdr = next(r for r in rows if r and r[0] == "Line No")
iL, iSrc, iAddr, iS = hdr.index("Line No"), 1, hdr.index("Address"), hdr.index("# Samples")
lines = [(int(r[iL]), r[iSrc].strip(), int(r[iS])) for r in rows if len(r) > iS and r[iAddr] == "-" and r[iL].isdigit()]
tot = sum(l[2] for l in lines)
print(f"samples {tot}")
src = [l[1] for l in sorted(lines)]          # the report carries the source it was built from
lineno = [l[0] for l in sorted(lines)]


def region_of(n):
    """the enclosing function / phase, from markers in the source"""
    k = lineno.index(n)
    for i in range(k, -1, -1):
        t = src[i]
        if "void process_batch(" in t: return "process_batch (refinement)"
        if "void consume_results(" in t: return "consume_results (fold)"
        if "push_phase:" in t: return "run_steps: push phase + trigger check"
        if "bool run_steps(" in t: return "run_steps: kick + drift"
        if "---- resume:" in t: return "kernel: resume of a handed-over tile"
        if "---- first segment:" in t: return "kernel: input staging + prologue + corrector"
        if "---- integration of steps" in t: return "kernel: step-loop driver / drain"
        if "---- hand the tile over" in t: return "kernel: hand-over"
        if "---- epilogue" in t: return "kernel: epilogue"
        if "ttvb_main_kernel(const" in t: return "kernel: setup"
    return "other"


regions = {}
for n, s, smp in lines:
    r = region_of(n)
    regions[r] = regions.get(r, 0) + smp
for k, v in sorted(regions.items(), key=lambda kv: -kv[1]):
    print(f"  {100 * v / tot:5.1f} %  {k}")
print("-- top lines")
for n, s, smp in sorted(lines, key=lambda l: -l[2])[:top]:
    print(f"  {100 * smp / tot:5.1f} %  {n:4d}  {s[:110]}")
