import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    from nauyaca_b200 import _lib
    _lib.ensure_built()          # nvcc cross-compiles without a GPU; a fresh checkout has no libttvb.so


def has_gpu():
    try:
        from nauyaca_b200 import _lib
        return _lib.lib().ttvb_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def systems():
    with open(os.path.join(GOLDEN, "systems.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def known():
    with open(os.path.join(GOLDEN, "known_answers.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def plumbing():
    with open(os.path.join(GOLDEN, "plumbing.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle_py
    oracle_py.build()
    return oracle_py


@pytest.fixture(scope="session")
def harness():
    """TEST-ONLY host build of the kernel arithmetic (tests/host_harness)."""
    import ctypes as C
    d = os.path.join(ROOT, "tests", "host_harness")
    so, src = os.path.join(d, "libhh.so"), os.path.join(d, "harness.cpp")
    hdr = os.path.join(ROOT, "nauyaca_b200", "csrc", "ttvb_math.cuh")
    if (not os.path.exists(so)) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-mfma", "-fPIC", "-shared", "-o", so, src])
    lib = C.CDLL(so)
    from oracle import oracle_py
    rcp, rsq = oracle_py.seed_tables()
    lib.hh_set_seed_tables(rcp.ctypes.data_as(C.c_void_p), rsq.ctypes.data_as(C.c_void_p))

    def eph(flat, mstar, t0, dt, total, cap=5000):
        flat = np.ascontiguousarray(flat, np.float64)
        n = C.c_int(0)
        pl = np.zeros(cap, np.int32); ep = np.zeros(cap, np.int32)
        tm = np.zeros(cap); rs = np.zeros(cap); vs = np.zeros(cap)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        st = lib.hh_ephemeris(p(flat), C.c_int(flat.size // 7), C.c_double(mstar), C.c_double(t0), C.c_double(dt),
                              C.c_double(total), C.c_int(cap), C.byref(n), p(pl), p(ep), p(tm), p(rs), p(vs))
        k = n.value
        return st, pl[:k], ep[:k], tm[:k], rs[:k], vs[:k]
    return eph


def walkers_near(osys_or_spec, center_cube, n, scale, seed):
    rng = np.random.default_rng(seed)
    c = np.asarray(center_cube)
    return np.clip(c[None, :] + scale * rng.standard_normal((n, c.size)), 0.0, 1.0)


from nauyaca_b200.engine import SpecSystem as SpecPS   # stand-in for nauyaca.PlanetarySystem (see its docstring)


@pytest.fixture(scope="session")
def spec_ps():
    return SpecPS
