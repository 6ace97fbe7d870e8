#!/usr/bin/env python
"""Development tool: locate the step loop of ttvb_main_kernel<N> in the SASS of a built library
and count local-memory (spill) instructions inside it.  usage: sass_loop_spills.py lib.so N"""
import re, subprocess, sys
lib, N = sys.argv[1], sys.argv[2]
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(sass) if "Function :" in l and f"main_kernelILi{N}E" in l)
end = next((i for i in range(start + 1, len(sass)) if "Function :" in sass[i]), len(sass))
ins = []
for l in sass[start:end]:
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", l)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
addr_index = {a: i for i, (a, _) in enumerate(ins)}
best = None
for i, (a, t) in enumerate(ins):
    m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?(0x[0-9a-f]+)", t)
    if m:
        tgt = int(m.group(1), 16)
        if tgt < a and tgt in addr_index:
            j = addr_index[tgt]
            body = ins[j:i + 1]
            nd = sum(1 for _, x in body if x.split()[-0].startswith(("DFMA", "DMUL", "DADD")) or " DFMA" in x or " DMUL" in x)
            if best is None or nd > best[0]:
                best = (nd, j, i)
nd, j, i = best
body = ins[j:i + 1]
cnt = lambda k: sum(1 for _, x in body if re.search(r"(^|\s)" + k, x))
print(f"kernel<{N}>: {len(ins)} instrs; largest backward-branch body: {len(body)} instrs @ {ins[j][0]:#x}..{ins[i][0]:#x}, "
      f"DFMA {cnt('DFMA')} DMUL {cnt('DMUL')} DADD {cnt('DADD')} LDL {cnt('LDL')} STL {cnt('STL')} LDS {cnt('LDS')} STS {cnt('STS')} MUFU {cnt('MUFU')}")
