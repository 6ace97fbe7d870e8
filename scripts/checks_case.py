"""The cases a compute-sanitizer run would cover, for the SELF-CHECKING kernel build
(-DTTVB_CHECKS, nauyaca_b200/libttvb_checks.so; compute-sanitizer itself is closed on the GPU pool,
profiles/r2a_compute_sanitizer_closed.txt):

    TTVB_LIB=nauyaca_b200/libttvb_checks.so TTVB_NSEG=3 python scripts/checks_case.py

Case A: C4-shaped batch (3 planets, near-mode walkers) cut into 3 time segments, so that the
        inter-CTA hand-over (context save, per-tile flag, resume) runs.
Case B: identical walkers: every lane of a warp triggers in the same step, the event queue
        saturates in the middle of a push round (re-entry into the push phase).
Case C: table mode (ttvb_ephemeris) and the device-resident PT sampler for a few iterations.
Case D: thin launches (systems on the first few lanes of every warp: what small batches get), forced to
        5 lanes for a 3000-system batch and left to the launch planner for the small batches of B and C.
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
nA = int(os.environ.get("CHK_NA", "40000"))     # more than one wave of tiles: units are handed between CTAs
XA = rng.random((nA, spec["ndim"]))
chi2, logl, st = eng.loglike(XA, mode=0)
oA = osys.loglike_threads(XA[:2000], mode=0)
assert np.array_equal(chi2[:2000], oA[0]) and np.array_equal(st[:2000], oA[2]), "case A mismatch"
sub = rng.permutation(nA)[:3000]
assert np.array_equal(eng.loglike(XA[sub], mode=0)[0], chi2[sub]), "case A: position dependence"
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
os.environ["TTVB_LANES"] = "5"
e3 = nb.BatchEngine(spec)
chi2, logl, st = e3.loglike(XA[:3000], mode=0)
assert np.array_equal(chi2[:2000], oA[0]) and np.array_equal(st[:2000], oA[2]), "case D mismatch"
del os.environ["TTVB_LANES"]
import ctypes  # noqa: E402
from nauyaca_b200 import _lib  # noqa: E402
counts = (ctypes.c_int32 * 8)()
rc = _lib.lib().ttvb_check_failures(0, counts)
print(json.dumps({"checks_case": "ok", "launches": eng.launch_count() + e2.launch_count() + e3.launch_count(), "check_rc": rc,
                  "violations": list(counts), "nseg": os.environ.get("TTVB_NSEG"), "lib": os.path.basename(_lib.LIB_PATH)}))
