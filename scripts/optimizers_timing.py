#!/usr/bin/env python
"""Timing of the batched optimizer chains against the notebook record of the reference
(Documentation/3- Running Optimizers.ipynb: 250 solutions in 6.933 min on 8 cores)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import nauyaca_b200 as nb
nsols = int(sys.argv[1]) if len(sys.argv) > 1 else 250
spec = json.load(open(os.path.join(ROOT, "tests", "golden", "systems.json")))["doc2"]
PS = nb.SpecSystem(spec)
t = time.time()
out = nb.optimizers.run_optimizers(PS, nsols=nsols, seed=1)
dt = time.time() - t
b = out["info"]["batches"]
print(json.dumps({"nsols": nsols, "seconds": dt, "launches": len(b), "evals": int(np.sum(b)), "median_batch": float(np.median(b)),
                  "best_chi2": float(out["chi2"][0]), "median_chi2": float(np.median(out["chi2"])),
                  "reference_record": "250 solutions in 6.933 min with cores=8 (notebook)"}))
