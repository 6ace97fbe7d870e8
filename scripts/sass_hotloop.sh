#!/bin/bash
# Development tool: compile the 2-/3-planet kernels only and print, for the step loop of the N-planet
# kernel ($1 = 2 or 3), the loop's span and its forward branches (scripts/sass_loops.py style).
set -e -o pipefail
N=${1:-2}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p "$ROOT/build"
cd /tmp
rm -f /tmp/t23.so
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -std=c++17 -lineinfo -shared -Xcompiler -fPIC -DTTVB_DEV_ONLY_23 ${TTVB_EXTRA} -Xptxas -v -o /tmp/t23.so "$ROOT/nauyaca_b200/csrc/ttvb.cu" 2>&1 | grep -B1 "spill" | grep -A1 "run_steps\|process_batch" | grep -v "^--" | paste - - | sed 's/ptxas info    : Function properties for //' | awk '{print $1, $2, $3, $4, $5, $6, $7, $8, $9, $10, $11, $12, $13, $14}'
cuobjdump -sass /tmp/t23.so | awk -v n="$N" '/Function : /{f = ($0 ~ ("ttvb_main_kernelILi" n "ELb0"))} f' > "$ROOT/build/k${N}b.sass"
python - "$N" "$ROOT" <<'PY'
import re, sys
n, root = sys.argv[1], sys.argv[2]
ins = []
for line in open(f'{root}/build/k{n}b.sass'):
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', line)
    if m: ins.append((int(m.group(1), 16), m.group(2).strip()))
def tgt(s):
    m = re.search(r'\bBRA(?:\.[A-Z.]+)?\s+(?:[!\w]+,\s*)*`?\(?\.?(0x[0-9a-f]+)', s)
    return int(m.group(1), 16) if m else None
loops = [(t, a, s) for a, s in ins for t in [tgt(s)] if t is not None and t <= a]
# the step loop: the LAST loop of more than 600 instructions whose back-edge is conditional
cand = [(t, a, s) for t, a, s in loops if (a - t) // 16 > 600 and s.startswith('@')]
t, a, s = cand[-1]
print(f"{len(ins)} instructions; step loop {t:#x}..{a:#x}: {(a - t) // 16} instructions  [{s}]")
for b, s2 in ins:
    if t <= b <= a:
        t2 = tgt(s2)
        if t2 is not None and (t2 - b > 16 * 30 or t2 < t) and 'DIV' not in s2:
            print(f"    {(b - t) // 16:5d} -> {(t2 - t) // 16:6d}  {s2}")
ops = [s2.split()[1] if s2.startswith('@') else s2.split()[0] for b, s2 in ins if t <= b <= a]
print("    LDS", sum(o.startswith('LDS') for o in ops), "STS", sum(o.startswith('STS') for o in ops), "REDUX", sum(o.startswith('REDUX') for o in ops))
PY
