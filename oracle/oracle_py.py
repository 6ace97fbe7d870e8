"""ctypes front-end of the CPU oracle (oracle/ttv_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs. The product package nauyaca_b200 never imports it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libttv_oracle.so")
_LIB_NATIVE = os.path.join(_HERE, "libttv_oracle_native.so")     # timing build (bench.py CPU legs)

ST_OUT_OF_BOUNDS, ST_KEPLER_FAIL, ST_BAD_TRANSIT, ST_EVENT_OVERFLOW, ST_INVALID_INPUT = 1, 2, 4, 8, 16


def build(force=False):
    src = os.path.join(_HERE, "ttv_oracle.c")
    for so in (_LIB, _LIB_NATIVE):
        if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", _HERE, "-B", os.path.basename(so)])
    return _LIB


class _Sys(C.Structure):
    _fields_ = [("npla", C.c_int), ("ndim", C.c_int), ("nconst", C.c_int), ("nobs", C.c_int),
                ("mstar", C.c_double), ("t0", C.c_double), ("ftime", C.c_double), ("dt", C.c_double),
                ("rstar_au", C.c_double), ("second_term", C.c_double),
                ("bi", C.c_void_p), ("bf", C.c_void_p), ("lo", C.c_void_p), ("hi", C.c_void_p),
                ("const_idx", C.c_void_p), ("const_val", C.c_void_p),
                ("obs_planet", C.c_void_p), ("obs_epoch", C.c_void_p),
                ("obs_time", C.c_void_p), ("obs_sigma", C.c_void_p)]


_lib = None
_lib_native = None
_lib_lock = __import__("threading").Lock()


def lib(native=False):
    with _lib_lock:
        return _lib_locked(native)


def _lib_locked(native=False):
    """The oracle library. native=True: the timing build (host divide/sqrt instead of the
    emulated GPU seeds; last-bit differences) that bench.py's CPU legs time - never the checker."""
    global _lib, _lib_native
    if native:
        if _lib_native is None:
            if not os.path.exists(_LIB_NATIVE):
                build()
            _lib_native = C.CDLL(_LIB_NATIVE)
        return _lib_native
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        _lib = C.CDLL(_LIB)
        _lib.ttvo_nsteps.restype = C.c_long
        _lib.ttvo_nsteps.argtypes = [C.c_double] * 3
        _lib.ttvo_chi2.restype = C.c_double
        _lib.ttvo_test_cbrt.restype = C.c_double
        _lib.ttvo_test_cbrt.argtypes = [C.c_double]
        rcp, rsq = seed_tables()
        _lib.ttvo_set_seed_tables(_p(rcp), _p(rsq))
    return _lib


_TABLES = None


def seed_tables():
    """High words of the sm_100a FP64 reciprocal / reciprocal-square-root seeds (MUFU.RCP64H over
    operands 1.m, MUFU.RSQ64H over 1.m and 2*1.m; m = 20 mantissa bits), stored as deltas in
    oracle/mufu_sm100a.npz (dumped by scripts/mufu_dump.cu; the sha256 of the raw tables is in
    the file). Returns (rcp[2^20], rsq[2^21]) uint32; the arrays stay alive for the C side."""
    global _TABLES
    if _TABLES is None:
        import hashlib
        z = np.load(os.path.join(_HERE, "mufu_sm100a.npz"))
        out = []
        for k in ("rcp", "rsq"):
            first = np.int64(z[k + "_first"])
            t = np.concatenate([[first], first + np.cumsum(z[k + "_delta"].astype(np.int64))]).astype(np.uint32)
            out.append(np.ascontiguousarray(t))
        assert [hashlib.sha256(t.tobytes()).hexdigest() for t in out] == list(z["sha256"])
        _TABLES = tuple(out)
    return _TABLES


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def nsteps(t0, total, dt):
    return int(lib().ttvo_nsteps(float(t0), float(total), float(dt)))


def ephemeris(flat, mstar, t0, dt, total, cap=5000, native=False):
    """Engine seam S1: returns (status, planet[int], epoch[int], time, rsky, vsky)."""
    flat = np.ascontiguousarray(flat, dtype=np.float64)
    npla = flat.size // 7
    n = C.c_int(0)
    pl = np.zeros(cap, np.int32); ep = np.zeros(cap, np.int32)
    tm = np.zeros(cap); rs = np.zeros(cap); vs = np.zeros(cap)
    st = lib(native).ttvo_ephemeris(_p(flat), C.c_int(npla), C.c_double(mstar), C.c_double(t0),
                              C.c_double(dt), C.c_double(total), C.c_int(cap), C.byref(n),
                              _p(pl), _p(ep), _p(tm), _p(rs), _p(vs))
    k = n.value
    return st, pl[:k].copy(), ep[:k].copy(), tm[:k].copy(), rs[:k].copy(), vs[:k].copy()


class OracleSystem:
    """Frozen per-system constants (what planetarysystem.py:184-350 derives)."""

    def __init__(self, spec, native=False):
        """spec: dict with the keys of tests/golden/systems.json entries (or the
        equivalent built from a PlanetarySystem by nauyaca_b200.system_spec).
        native=True: loglike() runs the timing build (see lib())."""
        self.spec = spec
        self.native = bool(native)
        self.npla, self.ndim = int(spec["npla"]), int(spec["ndim"])
        self.bi = np.array(spec["bi"], np.float64); self.bf = np.array(spec["bf"], np.float64)
        self.hyper_lo = np.array([h[0] for h in spec["hypercube"]], np.float64)
        self.hyper_hi = np.array([h[1] for h in spec["hypercube"]], np.float64)
        self.b_lo = np.array([h[0] for h in spec["bounds"]], np.float64)
        self.b_hi = np.array([h[1] for h in spec["bounds"]], np.float64)
        ck = sorted(int(k) for k in spec["constant_params"])
        self.const_idx = np.array(ck, np.int32)
        self.const_val = np.array([spec["constant_params"][str(k)] for k in ck], np.float64)
        op, oe, ot, os_ = [], [], [], []
        for pid, o in spec["observed"].items():
            op += [o["planet_index"]] * len(o["epochs"]); oe += o["epochs"]
            ot += o["times"]; os_ += o["sigma"]
        self.obs_planet = np.array(op, np.int32); self.obs_epoch = np.array(oe, np.int32)
        self.obs_time = np.array(ot, np.float64); self.obs_sigma = np.array(os_, np.float64)
        self.second_term = float(spec["second_term_sum"])

    def _struct(self, mode):
        s = _Sys()
        sp = self.spec
        s.npla, s.ndim, s.nconst, s.nobs = self.npla, self.ndim, len(self.const_idx), len(self.obs_time)
        s.mstar, s.t0, s.ftime, s.dt = sp["mstar"], sp["t0"], sp["ftime"], sp["dt"]
        s.rstar_au, s.second_term = sp["rstarAU"], self.second_term
        s.bi, s.bf = _p(self.bi), _p(self.bf)
        if mode == 0:
            s.lo, s.hi = _p(self.hyper_lo), _p(self.hyper_hi)
        else:
            s.lo, s.hi = _p(self.b_lo), _p(self.b_hi)
        s.const_idx, s.const_val = _p(self.const_idx), _p(self.const_val)
        s.obs_planet, s.obs_epoch = _p(self.obs_planet), _p(self.obs_epoch)
        s.obs_time, s.obs_sigma = _p(self.obs_time), _p(self.obs_sigma)
        return s

    def loglike(self, X, mode=0):
        """Returns (chi2[B], logl[B], status[B]) for X[B, ndim] (mode 0/1) or X[B, 7*npla] (mode 2)."""
        X = np.ascontiguousarray(np.atleast_2d(X), dtype=np.float64)
        B = X.shape[0]
        chi2 = np.zeros(B); logl = np.zeros(B); st = np.zeros(B, np.int32)
        s = self._struct(mode)
        lib(self.native).ttvo_loglike_batch(C.byref(s), _p(X), C.c_long(B), C.c_int(mode), _p(chi2), _p(logl), _p(st))
        return chi2, logl, st

    def loglike_threads(self, X, mode=0, threads=None):
        """loglike() with the rows spread over host threads (the C call releases the GIL): for
        samples of long integrations (BASELINE config 4: about 10 ms per system)."""
        from concurrent.futures import ThreadPoolExecutor
        X = np.ascontiguousarray(np.atleast_2d(X), dtype=np.float64)
        threads = threads or min(32, os.cpu_count() or 1)
        parts = [p for p in np.array_split(np.arange(X.shape[0]), threads * 4) if p.size]
        with ThreadPoolExecutor(threads) as ex:
            res = list(ex.map(lambda idx: self.loglike(X[idx], mode), parts))
        return tuple(np.concatenate([r[k] for r in res]) for k in range(3))

    def cube_to_physical(self, x):
        x = np.ascontiguousarray(x, np.float64)
        flat = np.zeros(7 * self.npla)
        lib().ttvo_cube_to_physical(_p(x), C.c_int(self.ndim), _p(self.bi), _p(self.bf), _p(self.const_idx),
                                    _p(self.const_val), C.c_int(len(self.const_idx)), C.c_int(self.npla), _p(flat))
        return flat

    def chi2_from_events(self, pl, ep, tm, rs):
        per = np.zeros(self.npla)
        pl = np.ascontiguousarray(pl, np.int32); ep = np.ascontiguousarray(ep, np.int32)
        tm = np.ascontiguousarray(tm, np.float64); rs = np.ascontiguousarray(rs, np.float64)
        tot = lib().ttvo_chi2(C.c_int(len(tm)), _p(pl), _p(ep), _p(tm), _p(rs), C.c_double(self.spec["rstarAU"]),
                              C.c_int(len(self.obs_time)), _p(self.obs_planet), _p(self.obs_epoch),
                              _p(self.obs_time), _p(self.obs_sigma), _p(per))
        return float(tot), per
