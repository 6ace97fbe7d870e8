#!/usr/bin/env python
"""Development tool: where the warp-stall samples of ttvb_main_kernel go.

    ncu -i rep.ncu-rep --page source --csv --print-source sass > sass.csv
    python scripts/profile_regions.py sass.csv <warp-steps per launch>

Every SASS instruction appears once in this view (the cuda,sass view repeats an instruction
under every line of its inline stack). Instructions are bucketed by how often they execute per
warp-step: ~1 = the step loop; the rest is the event path (push, refinement, fold), the rare
Kepler iterations and prologue/epilogue. Within the step loop the samples are split by pipe."""
import collections
import csv
import re
import sys

path, W = sys.argv[1], float(sys.argv[2])
rows = list(csv.reader(open(path)))
hdr, data = rows[1], rows[2:]
iS, iE, iSrc = hdr.index('# Samples'), hdr.index('Instructions Executed'), hdr.index('Source')
seq = [(r[iSrc].strip(), int(r[iS]), int(r[iE])) for r in data]
tot = sum(s[1] for s in seq)
print(f'kernel: {hdr and rows[0][1]}')
print(f'samples {tot}, SASS instructions {len(seq)}, warp-steps {W:.0f}')


def opcode(s):
    return re.sub(r'^@!?U?P\d+\s+', '', s).split()[0].split('.')[0]


print('-- by executions per warp-step')
b, n = collections.Counter(), collections.Counter()
for s, smp, exe in seq:
    r = exe / W
    key = '0.85-1.3 (step loop)' if 0.85 < r < 1.3 else ('<0.05' if r < 0.05 else ('0.05-0.85' if r < 0.85 else '>1.3'))
    b[key] += smp; n[key] += 1
for k in ('0.85-1.3 (step loop)', '>1.3', '0.05-0.85', '<0.05'):
    print(f'  {k:22s} {n[k]:6d} instructions  {100 * b[k] / tot:5.1f} % of samples')
main = [(s, smp) for s, smp, exe in seq if 0.85 < exe / W < 1.3]
FP64 = ('DFMA', 'DMUL', 'DADD', 'DSETP', 'DMNMX')
fp = [x for x in main if opcode(x[0]) in FP64]
print(f'-- step loop: {len(main)} instructions per warp-step, {len(fp)} on the FP64 pipe '
      f'({100 * sum(x[1] for x in fp) / tot:.1f} % of samples), {len(main) - len(fp)} others '
      f'({100 * sum(x[1] for x in main if opcode(x[0]) not in FP64) / tot:.1f} %)')
c, m = collections.Counter(), collections.Counter()
for s, smp in main:
    if opcode(s) not in FP64:
        c[opcode(s)] += smp; m[opcode(s)] += 1
print('   others by opcode: ' + ', '.join(f'{op} x{m[op]} {100 * v / tot:.2f}%' for op, v in c.most_common(10)))
