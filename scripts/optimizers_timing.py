#!/usr/bin/env python
"""Timing of the batched optimizer chains against the notebook record of the reference
(Documentation/3- Running Optimizers.ipynb: 250 solutions in 6.933 min on 8 cores).

    python scripts/optimizers_timing.py [nsols] [lockstep|scipy|both]

Prints one JSON line per method: wall time, launches, batch sizes, and the distribution of the final
chi2 (the two methods search differently, so their results are compared as distributions; every
reported chi2 is re-evaluated and must equal the engine's value for the returned vector)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import nauyaca_b200 as nb
nsols = int(sys.argv[1]) if len(sys.argv) > 1 else 250
which = sys.argv[2] if len(sys.argv) > 2 else "lockstep"
spec = json.load(open(os.path.join(ROOT, "tests", "golden", "systems.json")))["doc2"]
PS = nb.SpecSystem(spec)
eng = nb.BatchEngine(spec)
for method in (["lockstep", "scipy"] if which == "both" else [which]):
    t = time.time()
    out = nb.optimizers.run_optimizers(PS, nsols=nsols, seed=1, method=method)
    dt = time.time() - t
    b = out["info"]["batches"]
    c = np.sort(out["chi2"])
    again = eng.loglike(out["cube"], mode=0)[0]
    q = lambda p: float(np.quantile(c, p))
    print(json.dumps({"method": method, "nsols": nsols, "seconds": dt, "launches": len(b), "evals": int(np.sum(b)),
                      "median_batch": float(np.median(b)), "best_chi2": float(c[0]),
                      "chi2_quantiles_10_25_50_75_90": [q(0.1), q(0.25), q(0.5), q(0.75), q(0.9)],
                      "below_100": int((c < 100).sum()), "below_1000": int((c < 1000).sum()),
                      "reported_equals_reevaluated": bool(np.array_equal(again, out["chi2"])),
                      "reference_record": "250 solutions in 6.933 min with cores=8 (notebook)"}))
