#!/usr/bin/env python
"""Development tool: the same launches again and again - C4 (time-segmented, full GPU), the one-system
ephemeris that builds its observed table, C5 chunk, a thin C3 launch - and the number of DISTINCT result
hashes per shape (must be 1: a race would show up as a second one)."""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
import nauyaca_b200 as nb  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 25
h = lambda *a: hashlib.sha1(b"".join(np.ascontiguousarray(x).tobytes() for x in a)).hexdigest()[:12]
out = {}
obs_hashes = set()
for r in range(4):
    spec, X, desc = bench.make_workload("c4", lambda s: nb.BatchEngine(s))
    obs_hashes.add(hashlib.sha1(json.dumps(spec["observed"], sort_keys=True).encode()).hexdigest()[:12])
out["c4 observed table (one-system ephemeris, thin launch), 4 builds"] = sorted(obs_hashes)
for wl, B in (("c4", 100000), ("c5", 300000), ("c3", 5000), ("c2", 10000)):
    spec, X, desc = bench.make_workload(wl, lambda s: nb.BatchEngine(s))
    X = X[:B]
    eng = nb.BatchEngine(spec)
    Xd = torch.from_numpy(X).cuda()
    seen = {}
    for i in range(reps):
        chi2, logl, st = eng.loglike(Xd)
        torch.cuda.synchronize()
        k = h(chi2.cpu().numpy(), st.cpu().numpy())
        seen[k] = seen.get(k, 0) + 1
    c2, l2, s2 = eng.loglike(X)                      # host-pointer path
    seen_host = h(c2, s2)
    out[f"{wl} B={B} x{reps}"] = {"device": seen, "host": seen_host}
    eng.close()
print(json.dumps(out, indent=1))
bad = [k for k, v in out.items() if (isinstance(v, list) and len(v) != 1) or (isinstance(v, dict) and (len(v["device"]) != 1 or v["host"] not in v["device"]))]
print("DETERMINISTIC" if not bad else f"NOT DETERMINISTIC: {bad}")
