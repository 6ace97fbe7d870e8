"""Small invocations of the hot path for compute-sanitizer (memcheck / racecheck / synccheck).

    TTVB_NSEG=3 compute-sanitizer --tool racecheck python scripts/sanitize_case.py

Case A: C4-shaped batch (3 planets, near-mode walkers) cut into 3 time segments, so that the
        inter-CTA hand-over (context save, per-tile flag, resume) runs.
Case B: identical walkers: every lane of a warp triggers in the same step, the event queue
        saturates in the middle of a push round (re-entry into the push phase).
Case C: table mode (ttvb_ephemeris) and the device-resident PT sampler for a few iterations.
Results are compared with the CPU oracle (bit equality), so a sanitizer run is also a parity run.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nauyaca_b200 as nb                      # noqa: E402
from oracle import oracle_py                   # noqa: E402  (checker only)

with open(os.path.join(ROOT, "tests", "golden", "systems.json")) as fh:
    systems = json.load(fh)
spec = systems["sys3"]
osys = oracle_py.OracleSystem(spec)
eng = nb.BatchEngine(spec)
rng = np.random.default_rng(3)
nA = int(os.environ.get("SAN_NA", "640"))
XA = rng.random((nA, spec["ndim"]))
chi2, logl, st = eng.loglike(XA, mode=0)
o = osys.loglike(XA, mode=0)
assert np.array_equal(chi2, o[0]) and np.array_equal(st, o[2]), "case A mismatch"
XB = np.repeat(rng.random((1, spec["ndim"])), 256, axis=0)
chi2, logl, st = eng.loglike(XB, mode=0)
o = osys.loglike(XB[:1], mode=0)
assert np.all(chi2 == o[0][0]), "case B mismatch"
flat = osys.cube_to_physical(XA[0])
n, pl, ep, tm, rs, vs, s2 = eng.ephemeris(np.stack([flat] * 40), mode=2)
oe = oracle_py.ephemeris(flat, spec["mstar"], spec["t0"], spec["dt"], spec["ftime"])
assert np.all(n == len(oe[3])) and np.array_equal(tm[7, :n[7]], oe[3]), "case C mismatch"
from nauyaca_b200.ptsampler import DeviceSampler   # noqa: E402
doc = systems["doc2"]
e2 = nb.BatchEngine(doc)
smp = DeviceSampler(e2, nwalkers=32, ntemps=3, seed=5)
p0 = 0.45 + 0.1 * rng.random((3, 32, doc["ndim"]))
for _ in smp.sample(p0, iterations=3, adapt=True):
    pass
print("sanitize_case ok: launches", eng.launch_count() + e2.launch_count())
