#!/usr/bin/env python
"""Development tool: kernel time for synthetic systems of N planets (all parameters free)."""
import os, sys, types
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
src = open(os.path.join(ROOT, "tests", "test_gpu_robustness.py")).read().replace("pytestmark = pytest.mark.gpu", "")
m = types.ModuleType("r"); exec(compile(src, "r", "exec"), m.__dict__)
import torch
import nauyaca_b200 as nb
for npla in [int(a) for a in sys.argv[1:]] or [1, 2, 3, 4, 5, 6]:
    spec = m._spec(npla, ftime=200.0, dt=0.05)
    eng = nb.BatchEngine(spec)
    X = torch.from_numpy(np.random.default_rng(0).random((40000, spec["ndim"]))).cuda()
    ms = []
    for _ in range(3):
        c, l, s = eng.loglike(X); torch.cuda.synchronize(); ms.append(eng.last_kernel_ms())
    steps = eng.nsteps
    print(f"N={npla} B=40000 nsteps={steps} kernel_ms {min(ms):8.2f}  planet-steps/s {40000 * steps * npla / (min(ms) * 1e-3):.3e}  lib {os.path.basename(os.environ.get('TTVB_LIB', 'libttvb.so'))}")
    eng.close()
