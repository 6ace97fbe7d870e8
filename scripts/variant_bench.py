#!/usr/bin/env python
"""Kernel-variant timing helper (development tool): times the resident-input C4/C2 batch with the
library given by $TTVB_LIB and checks the results against a reference library's bits.

    TTVB_LIB=build/variants/libttvb_x.so python scripts/variant_bench.py [c4|c2] [B]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
import nauyaca_b200 as nb  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "c4"
spec, X, desc = bench.make_workload(wl, lambda s: nb.BatchEngine(s))
if len(sys.argv) > 2:
    X = X[: int(sys.argv[2])]
eng = nb.BatchEngine(spec)
Xd = torch.from_numpy(X).cuda()
ms = []
for i in range(4):
    chi2, logl, st = eng.loglike(Xd)
    torch.cuda.synchronize()
    ms.append(eng.last_kernel_ms())
h = float(np.sum(chi2.cpu().numpy()[np.isfinite(chi2.cpu().numpy())]))
import hashlib
dig = hashlib.sha1(chi2.cpu().numpy().tobytes()).hexdigest()[:12]
print(f"{os.path.basename(os.environ.get('TTVB_LIB', 'libttvb.so')):40s} {wl} B={X.shape[0]} kernel_ms min {min(ms[1:]):9.3f} "
      f"evals/s {X.shape[0] / (min(ms[1:]) * 1e-3):12.0f} chi2sha {dig}")
