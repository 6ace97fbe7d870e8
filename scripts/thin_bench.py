#!/usr/bin/env python
"""Thin-launch timing helper (development tool): the latency-bound shapes (C2, C3, optimizer-sized
batches) with systems on the first L lanes of every warp only (TTVB_LANES=L), bits compared across L.

    python scripts/thin_bench.py
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
import nauyaca_b200 as nb  # noqa: E402

cases = [("c3", 5000), ("c2", 10000), ("c3", 122), ("c3", 2500), ("c3", 10000), ("c3", 18000), ("c5", 5000), ("c5", 15000)]
for wl, B in cases:
    spec, X, desc = bench.make_workload(wl, lambda s: nb.BatchEngine(s))
    X = np.tile(X, ((B + len(X) - 1) // len(X), 1))[:B]
    Xd = torch.from_numpy(X).cuda()
    ref = None
    for lanes in ("auto", "32", "24", "17", "16", "12", "9", "8", "4", "1"):
        if lanes == "auto":
            os.environ.pop("TTVB_LANES", None)
        else:
            os.environ["TTVB_LANES"] = lanes
            if (B + 4 * int(lanes) - 1) // (4 * int(lanes)) > 296:      # more CTAs than are resident: pointless
                continue
        eng = nb.BatchEngine(spec)
        ms = []
        for i in range(5):
            chi2, logl, st = eng.loglike(Xd)
            torch.cuda.synchronize()
            ms.append(eng.last_kernel_ms())
        dig = hashlib.sha1(chi2.cpu().numpy().tobytes() + st.cpu().numpy().tobytes()).hexdigest()[:12]
        ref = ref or dig
        ctas = "" if lanes == "auto" else f"ctas={(B + 4 * int(lanes) - 1) // (4 * int(lanes)):4d}"
        print(f"{wl} B={B:6d} lanes={lanes:>4s} {ctas:9s} kernel_ms {min(ms[1:]):8.3f}  {'same bits' if dig == ref else 'BITS DIFFER'}",
              flush=True)
        eng.close()
