#!/usr/bin/env python
"""Development tool: where the warp-stall samples of ttvb_main_kernel go.

    ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass > src.csv
    python scripts/profile_regions.py src.csv <warp-steps per launch> [top]

Splits the SASS instructions into those executed about once per warp-step (the step loop) and
the rest (event push, refinement, fold), and lists the rest by source line."""
import collections
import csv
import sys

path, warpsteps = sys.argv[1], float(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
rows = list(csv.reader(open(path)))
cur = hdr = line = None
agg = {}
for r in rows:
    if r and r[0] == 'File Path':
        cur = r[1].split('/')[-1]; continue
    if r and r[0] == 'Line No':
        hdr = r; iS = hdr.index('# Samples'); iE = hdr.index('Instructions Executed'); continue
    if not hdr or not r or r[0] == 'Function Name':
        continue
    if r[0] != '':
        line = (cur, int(r[0]), r[1].strip()[:90]); continue
    try:
        smp, exe = int(r[iS]), int(r[iE])
    except ValueError:
        continue
    k = (line, exe > 0.85 * warpsteps)
    agg[k] = agg.get(k, 0) + smp
tot = sum(agg.values())
main = sum(v for (l, m), v in agg.items() if m)
print(f'samples {tot}; step-loop instructions {main / tot:.3f}; rest {1 - main / tot:.3f}')
c = collections.Counter()
for (l, m), v in agg.items():
    if not m:
        c[l[0]] += v
print('rest by file:', {k: round(v / tot, 4) for k, v in c.items() if v / tot > 0.001})
for m in (False, True):
    print('--- top lines,', 'step loop' if m else 'rest')
    for v, l in sorted([(v, l) for (l, mm), v in agg.items() if mm == m], reverse=True)[:top]:
        print(f'{v / tot * 100:5.2f}% {l[0]}:{l[1]} {l[2]}')
