#!/usr/bin/env python
"""Development tool: in-order issue model of the hot path of the step loop.

Input: the CSV of `ncu -i rep --page source --csv` (SASS view). Takes the instructions that are
executed about once per warp-step (the common path through the loop, in address order),
parses their register operands and replays them through a simple in-order, single-warp
issue model with fixed latencies (DFMA/DMUL/DADD 8 cycles and a 2-cycle issue interval on the
FP64 pipe as measured by scripts/microbench/fp64_latency.cu, MUFU 19, LDS 29, ALU 4.5).
Prints the modelled cycles per step, the critical-path length, and where the stall cycles go."""
import csv, re, sys, collections

path, nwarpsteps = sys.argv[1], float(sys.argv[2])
rows = list(csv.reader(open(path)))
hdr, data = rows[1], rows[2:]
isrc, iexe = hdr.index('Source'), hdr.index('Instructions Executed')
hot = [r[isrc].strip() for r in data if 0.85 < int(r[iexe]) / nwarpsteps < 1.15]
LAT = {'DFMA': 8, 'DMUL': 8, 'DADD': 8, 'DSETP': 10, 'MUFU': 19, 'LDS': 29, 'LDC': 12, 'LDCU': 12, 'LD': 40, 'LDG': 300}
FP64 = ('DFMA', 'DMUL', 'DADD', 'DSETP')

def regs(tok):
    out = []
    for m in re.finditer(r'\b(U?R|U?P)(\d+)\b', tok):
        out.append(m.group(1) + m.group(2))
    return out

ready = collections.defaultdict(float)
t = 0.0
fp64_free = 0.0
stall_by_op = collections.Counter()
crit = collections.defaultdict(float)     # dependency-chain depth in cycles
n = 0
for s in hot:
    s2 = re.sub(r'^@!?U?P\d+\s+', '', s)
    pred = re.match(r'^@!?(U?P\d+)', s)
    parts = s2.split(None, 1)
    op = parts[0].split('.')[0]
    ops = parts[1] if len(parts) > 1 else ''
    toks = [x.strip() for x in ops.split(',')]
    wide = op in FP64 or op.startswith('D') or '.64' in parts[0]
    dst = regs(toks[0]) if toks and op not in ('STS', 'ST', 'STG', 'BRA', 'BSSY', 'BSYNC', 'EXIT', 'BAR', 'WARPSYNC', 'NOP') else []
    srcs = []
    for tk in (toks[1:] if dst else toks):
        srcs += regs(tk)
    if pred: srcs.append(pred.group(1))
    def expand(r):
        if r.startswith('R') and wide:
            k = int(r[1:]); return [f'R{k}', f'R{k+1}']
        return [r]
    src_ready = max([ready[x] for r in srcs for x in expand(r)] + [0.0])
    src_depth = max([crit[x] for r in srcs for x in expand(r)] + [0.0])
    issue = max(t + 1.0, src_ready)
    if op in FP64:
        issue = max(issue, fp64_free)
        fp64_free = issue + 2.0
    stall = issue - (t + 1.0)
    stall_by_op[op] += stall
    lat = LAT.get(op, 4.5)
    for r in dst:
        for x in expand(r):
            ready[x] = issue + lat
            crit[x] = src_depth + lat
    t = issue
    n += 1
print(f"hot-path instructions per step: {n}")
print(f"modelled single-warp cycles per step (in-order issue): {t:.0f}")
print(f"longest register dependency chain through the step: {max(crit.values()):.0f} cycles")
print(f"FP64-pipe floor (2 cycles per FP64 instruction): {2 * sum(1 for s in hot if re.sub(r'^@!?U?P[0-9]+ +', '', s).split('.')[0].split()[0] in FP64)}")
print("stall cycles by opcode of the stalled instruction:", [(k, round(v)) for k, v in stall_by_op.most_common(8)])
