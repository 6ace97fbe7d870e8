#!/usr/bin/env python
"""Development tool: kernel time of the Documentation two-planet system for several batch sizes
(latency-bound below one wave of 37 888 systems, throughput-bound above)."""
import os, sys, json, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nauyaca_b200 as nb
spec = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "systems.json")))["doc2"]
eng = nb.BatchEngine(spec)
for B in (32, 592, 2368, 5000, 10000, 18944, 37888, 75776, 400000):
    X = np.random.default_rng(1).random((B, spec["ndim"]))
    Xd = torch.from_numpy(X).cuda()
    ms = []
    for i in range(4):
        chi2, logl, st = eng.loglike(Xd); torch.cuda.synchronize(); ms.append(eng.last_kernel_ms())
    import hashlib
    print(os.path.basename(os.environ.get("TTVB_LIB", "libttvb.so")), "B", B, "ms", round(min(ms[1:]), 3), "evals/s", int(B / (min(ms[1:]) * 1e-3)), hashlib.sha1(chi2.cpu().numpy().tobytes()).hexdigest()[:10])
