#!/usr/bin/env python
"""Golden vectors for nauyaca_b200.walkers.init_walkers: runs the UNMODIFIED reference
``nauyaca.utils.init_walkers`` (utils.py:463-716) from /root/reference on seeded inputs and
stores inputs and outputs in tests/golden/init_walkers.json.

Run once in the development container:  python tests/golden/make_init_walkers.py
"""
import contextlib
import io
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import _stub_modules, REF  # noqa: E402

_stub_modules()
sys.path.insert(0, REF)
from nauyaca import utils as ref_utils  # noqa: E402

ndim, nsols = 5, 23
rng = np.random.default_rng(4242)
cube = rng.random((nsols, ndim))
cube[:, 3] = 0.37                                     # a dimension with zero spread (sig == 0 branch)
chi2 = rng.random(nsols) * 100.0
chi2[[4, 11]] = 1e20                                  # two failed optimizer runs
hyper = [[0.0, 1.0], [0.1, 0.9], [0.0, 1.0], [0.0, 1.0], [0.2, 1.0]]
PS = types.SimpleNamespace(ndim=ndim, hypercube=hyper, bounds=[[0.0, 1.0]] * ndim)
cases = []
for k, (dist, use_dict, fbest, ntemps, nwalkers) in enumerate([
        ("Uniform", False, 1.0, 3, 12), ("Gaussian", True, 1.0, 2, 10), ("Gaussian", False, 0.5, 3, 10),
        ("Picked", True, 0.7, 2, 14), ("Picked", False, 1.0, 4, 10), ("Ladder", True, 1.0, 4, 10),
        ("Ladder", False, 0.6, 5, 12), ("ladder", True, 0.3, 3, 10)]):
    opt = {"chi2": chi2.copy(), "cube": cube.copy()} if use_dict else np.column_stack((chi2, cube))
    np.random.seed(1000 + k)
    with contextlib.redirect_stdout(io.StringIO()):
        p0 = ref_utils.init_walkers(PS, distribution=dist, opt_data=None if dist == "Uniform" else opt,
                                    ntemps=ntemps, nwalkers=nwalkers, fbest=fbest)
    cases.append({"distribution": dist, "opt_as_dict": use_dict, "fbest": fbest, "ntemps": ntemps,
                  "nwalkers": nwalkers, "seed": 1000 + k, "shape": list(p0.shape),
                  "p0_hex": [float(v).hex() for v in np.asarray(p0, dtype=float).ravel()]})
out = {"ndim": ndim, "hypercube": hyper, "chi2": [float(v).hex() for v in chi2],
       "cube_hex": [[float(v).hex() for v in row] for row in cube], "cases": cases}
with open(os.path.join(HERE, "init_walkers.json"), "w") as fh:
    json.dump(out, fh)
print("wrote", len(cases), "cases")
