#!/usr/bin/env python
"""Development tool: backward branches (loops) of a function in `cuobjdump -sass` output, with their
span in instructions and the forward branches taken over more than 32 instructions inside the largest
loop - how much cold code the hot path of the step loop has to jump over.

    cuobjdump -sass -fun <mangled name> lib.so | python scripts/sass_loops.py
"""
import re
import sys

ins = []
for line in sys.stdin:
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
print(f"{len(ins)} instructions")
loops = []
for a, s in ins:
    m = re.search(r'\bBRA(?:\.[A-Z.]+)?\s+(?:U?P?\w+,\s*)?`?\(?\.?(0x[0-9a-f]+)', s)
    if m:
        t = int(m.group(1), 16)
        if t <= a:
            loops.append(((a - t) // 16, t, a, s))
loops.sort(reverse=True)
for span, t, a, s in loops[:6]:
    print(f"loop {t:#x}..{a:#x}: {span} instructions   [{s}]")
if loops:
    span, t, a, s = loops[0]
    for b, s2 in ins:
        if t <= b <= a:
            m = re.search(r'\bBRA(?:\.[A-Z.]+)?\s+(?:U?P?\w+,\s*)?`?\(?\.?(0x[0-9a-f]+)', s2)
            if m and int(m.group(1), 16) > b + 32 * 16:
                print(f"   forward {b:#x} -> {int(m.group(1), 16):#x} over {(int(m.group(1), 16) - b) // 16:5d}: {s2}")
