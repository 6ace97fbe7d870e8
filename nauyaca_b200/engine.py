"""Batched TTV likelihood engine: Python face of the C ABI (include/ttvb.h).

``BatchEngine`` is built from a ``nauyaca.PlanetarySystem`` (or the equivalent dict, see
``system_spec``) and evaluates whole batches of parameter vectors on one B200. PyTorch is
used only to hand over device pointers (``tensor.data_ptr()``) and streams.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import (MODE_CUBE, MODE_PHYSICAL, MODE_PHYSICAL_FULL, EVENT_CAP, TTVBError)


def system_spec(PS):
    """Freeze the attributes of a nauyaca.PlanetarySystem that the hot path consumes
    (reference planetarysystem.py:184-350) into a plain dict. A dict with the same keys
    (e.g. an entry of tests/golden/systems.json) is returned unchanged."""
    if isinstance(PS, dict):
        return PS
    obs = {}
    for pid in PS.transit_times:
        obs[pid] = {"planet_index": int(PS.planets_IDs[pid]),
                    "epochs": [int(e) for e in PS.transit_times[pid].keys()],
                    "times": [float(t) for t in PS.transit_times[pid].values()],
                    "sigma": [float(PS.sigma_obs[pid][e]) for e in PS.transit_times[pid].keys()]}
    return {"system_name": getattr(PS, "system_name", ""),
            "mstar": float(PS.mstar), "rstarAU": float(PS.rstarAU),
            "t0": float(PS.t0), "ftime": float(PS.ftime), "dt": float(PS.dt),
            "npla": int(PS.npla), "ndim": int(PS.ndim),
            "planets_IDs": {k: int(v) for k, v in PS.planets_IDs.items()},
            "bounds": [[float(a), float(b)] for a, b in PS.bounds],
            "bi": [float(v) for v in PS.bi], "bf": [float(v) for v in PS.bf],
            "hypercube": [[float(a), float(b)] for a, b in PS.hypercube],
            "constant_params": {str(int(k)): float(v) for k, v in PS.constant_params.items()},
            "observed": obs,
            # Python's sum() over the dict values, as utils.py:397 does
            "second_term_sum": float(sum(PS.second_term_logL.values()))}


class SpecSystem:
    """Attribute container with the PlanetarySystem attributes the hot path reads
    (planetarysystem.py:184-350), rebuilt from a ``system_spec`` dict. Lets the functions that
    take a ``PSystem`` (utils.py mirror, GpuPool, make_sampler) run where the reference package
    is not importable."""

    def __init__(self, spec):
        from collections import OrderedDict
        self.system_name = spec.get("system_name", "")
        self.mstar, self.rstarAU = spec["mstar"], spec["rstarAU"]
        self.rstar = spec.get("rstar")
        self.t0, self.ftime, self.dt = spec["t0"], spec["ftime"], spec["dt"]
        self.npla, self.ndim = spec["npla"], spec["ndim"]
        self.planets_IDs = dict(spec.get("planets_IDs", {p: o["planet_index"] for p, o in spec["observed"].items()}))
        self.bounds = [list(b) for b in spec["bounds"]]
        self.bi, self.bf = list(spec["bi"]), list(spec["bf"])
        self.hypercube = [list(h) for h in spec["hypercube"]]
        self.constant_params = OrderedDict((int(k), v) for k, v in
                                           sorted(spec["constant_params"].items(), key=lambda kv: int(kv[0])))
        self.transit_times = {p: dict(zip(o["epochs"], o["times"])) for p, o in spec["observed"].items()}
        self.sigma_obs = {p: dict(zip(o["epochs"], o["sigma"])) for p, o in spec["observed"].items()}
        self.second_term_logL = dict(spec.get("second_term_logL", {"sum": spec["second_term_sum"]}))
        self.params_names = spec.get("params_names", "")
        self.params_names_all = spec.get("params_names_all", "")


def _is_torch(x):
    return type(x).__module__.startswith("torch")


class BatchEngine:
    """One planetary system resident on one GPU; every method evaluates a batch."""

    def __init__(self, PS, device=0):
        self.spec = spec = system_spec(PS)
        self.npla, self.ndim = int(spec["npla"]), int(spec["ndim"])
        self.device = int(device)
        f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        self._bi, self._bf = f64(spec["bi"]), f64(spec["bf"])
        self._clo = f64([h[0] for h in spec["hypercube"]]); self._chi = f64([h[1] for h in spec["hypercube"]])
        self._plo = f64([h[0] for h in spec["bounds"]]); self._phi = f64([h[1] for h in spec["bounds"]])
        ck = sorted(int(k) for k in spec["constant_params"])
        self._cidx = np.ascontiguousarray(ck, dtype=np.int32)
        self._cval = f64([spec["constant_params"][str(k)] for k in ck])
        op, oe, ot, os_ = [], [], [], []
        for _pid, o in spec["observed"].items():
            op += [o["planet_index"]] * len(o["epochs"]); oe += o["epochs"]; ot += o["times"]; os_ += o["sigma"]
        self._op = np.ascontiguousarray(op, dtype=np.int32); self._oe = np.ascontiguousarray(oe, dtype=np.int32)
        self._ot, self._os = f64(ot), f64(os_)
        self.nobs = len(ot)
        s = _lib.ttvb_system()
        s.npla, s.ndim, s.nconst, s.nobs = self.npla, self.ndim, len(ck), self.nobs
        s.mstar, s.t0, s.ftime, s.dt = spec["mstar"], spec["t0"], spec["ftime"], spec["dt"]
        s.rstar_au, s.second_term = spec["rstarAU"], spec["second_term_sum"]
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        s.bi, s.bf = p(self._bi), p(self._bf)
        s.cube_lo, s.cube_hi, s.phys_lo, s.phys_hi = p(self._clo), p(self._chi), p(self._plo), p(self._phi)
        s.const_idx, s.const_val = p(self._cidx), p(self._cval)
        s.obs_planet, s.obs_epoch, s.obs_time, s.obs_sigma = p(self._op), p(self._oe), p(self._ot), p(self._os)
        self._h = C.c_void_p()
        _lib.check(_lib.lib().ttvb_create(C.byref(s), self.device, C.byref(self._h)))
        n = C.c_int64()
        _lib.check(_lib.lib().ttvb_nsteps(self._h, C.byref(n)))
        self.nsteps = int(n.value)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            _lib.lib().ttvb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_cube_bounds(self, hypercube):
        """Follow an in-place change of PSystem.hypercube (optimizers.py:90)."""
        lo = np.ascontiguousarray([h[0] for h in hypercube], dtype=np.float64)
        hi = np.ascontiguousarray([h[1] for h in hypercube], dtype=np.float64)
        if lo.size != self.ndim:
            raise ValueError("hypercube must have ndim rows")
        _lib.check(_lib.lib().ttvb_set_cube_bounds(self._h, lo.ctypes.data, hi.ctypes.data))
        self._clo, self._chi = lo, hi

    # ---------------------------------------------------------------- helpers
    def rowlen(self, mode):
        return 7 * self.npla if mode == MODE_PHYSICAL_FULL else self.ndim

    def _prep_host(self, X, mode):
        X = np.ascontiguousarray(np.atleast_2d(np.asarray(X, dtype=np.float64)))
        if X.shape[1] != self.rowlen(mode):
            raise ValueError(f"expected rows of length {self.rowlen(mode)} for mode {mode}, got {X.shape[1]}")
        return X

    # ---------------------------------------------------------------- objective
    def loglike(self, X, mode=MODE_CUBE, per_planet=False, stream=None):
        """chi2, logl, status (and per-planet chi2) for X[B, rowlen].

        numpy in -> numpy out (host buffers, copies inside the call);
        CUDA torch tensors in -> torch tensors out (asynchronous on `stream`)."""
        L = _lib.lib()
        if _is_torch(X):
            import torch
            if not X.is_cuda:
                raise TTVBError("torch input must be a CUDA tensor (use numpy for host buffers)")
            if X.dtype != torch.float64 or not X.is_contiguous() or X.dim() != 2 or X.shape[1] != self.rowlen(mode):
                raise ValueError("X must be a contiguous float64 CUDA tensor [B, rowlen]")
            if X.device.index != self.device:
                raise ValueError(f"X lives on cuda:{X.device.index}, the engine on cuda:{self.device}")
            B = X.shape[0]
            chi2 = torch.empty(B, dtype=torch.float64, device=X.device)
            logl = torch.empty(B, dtype=torch.float64, device=X.device)
            st = torch.empty(B, dtype=torch.int32, device=X.device)
            pl = torch.empty((B, self.npla), dtype=torch.float64, device=X.device) if per_planet else None
            sp = stream if stream is not None else torch.cuda.current_stream(X.device).cuda_stream
            _lib.check(L.ttvb_loglike(self._h, X.data_ptr(), B, mode, chi2.data_ptr(), logl.data_ptr(),
                                      pl.data_ptr() if per_planet else None, st.data_ptr(), C.c_void_p(sp)))
            return (chi2, logl, st, pl) if per_planet else (chi2, logl, st)
        X = self._prep_host(X, mode)
        B = X.shape[0]
        chi2 = np.empty(B); logl = np.empty(B); st = np.empty(B, np.int32)
        pl = np.empty((B, self.npla)) if per_planet else None
        _lib.check(L.ttvb_loglike(self._h, X.ctypes.data, B, mode, chi2.ctypes.data, logl.ctypes.data,
                                  pl.ctypes.data if per_planet else None, st.ctypes.data,
                                  C.c_void_p(stream) if stream else None))
        return (chi2, logl, st, pl) if per_planet else (chi2, logl, st)

    def loglike_into(self, x_ptr, B, mode, chi2_ptr, logl_ptr, status_ptr, stream=None):
        """Raw-pointer form (host or device pointers, caller-owned buffers). Device pointers must
        belong to the engine's device (self.device): the kernel is launched there."""
        _lib.check(_lib.lib().ttvb_loglike(self._h, x_ptr, B, mode, chi2_ptr, logl_ptr, None, status_ptr,
                                           C.c_void_p(stream) if stream else None))

    def ephemeris(self, X, mode=MODE_PHYSICAL_FULL, cap=EVENT_CAP):
        """Event tables for X[B, rowlen] (host arrays): n_events[B], planet, epoch, time, rsky,
        vsky [B, cap], status[B]."""
        X = self._prep_host(X, mode)
        B = X.shape[0]
        n = np.zeros(B, np.int32); st = np.zeros(B, np.int32)
        pl = np.full((B, cap), -2, np.int32); ep = np.full((B, cap), -2, np.int32)
        tm = np.full((B, cap), -2.0); rs = np.full((B, cap), -2.0); vs = np.full((B, cap), -2.0)
        _lib.check(_lib.lib().ttvb_ephemeris(self._h, X.ctypes.data, B, mode, cap, n.ctypes.data, pl.ctypes.data,
                                             ep.ctypes.data, tm.ctypes.data, rs.ctypes.data, vs.ctypes.data,
                                             st.ctypes.data, None))
        # slots past n_events keep TTVFast's DEFAULT value -2 (utils.py:76-79 relies on it)
        idx = np.arange(cap)[None, :] >= n[:, None]
        for a in (pl, ep):
            a[idx] = -2
        for a in (tm, rs, vs):
            a[idx] = -2.0
        return n, pl, ep, tm, rs, vs, st

    def cube_to_physical(self, X):
        X = self._prep_host(X, MODE_CUBE)
        B = X.shape[0]
        flat = np.empty((B, 7 * self.npla)); inside = np.empty(B, np.int32)
        _lib.check(_lib.lib().ttvb_cube_to_physical(self._h, X.ctypes.data, B, flat.ctypes.data,
                                                    inside.ctypes.data, None))
        return flat, inside.astype(bool)

    # ---------------------------------------------------------------- measurement
    def last_kernel_ms(self):
        ms = C.c_float()
        _lib.check(_lib.lib().ttvb_last_kernel_ms(self._h, C.byref(ms)))
        return float(ms.value)

    def launch_count(self):
        n = C.c_int64()
        _lib.check(_lib.lib().ttvb_launch_count(self._h, C.byref(n)))
        return int(n.value)

    def stream_retries(self):
        """Streamed host-pointer launches that had to be run again (see ttvb_stream_retries)."""
        n = C.c_int64()
        _lib.check(_lib.lib().ttvb_stream_retries(self._h, C.byref(n)))
        return int(n.value)

    def fp64_peak_tflops(self):
        t = C.c_double(); ms = C.c_float()
        _lib.check(_lib.lib().ttvb_fp64_peak(self._h, C.byref(t), C.byref(ms)))
        return float(t.value)
