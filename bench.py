#!/usr/bin/env python
"""bench.py -- likelihood evaluations per second of the batched TTV engine on B200.

    python bench.py --gpus N --steps K --warmup W [--workload c4|c2|c5|c3] [--impl reference]

A "step" is one pass of the hot path (cube vector -> ephemerides -> chi2 -> logL) over one
batch of synthetic walkers. Default workload = BASELINE.json configs[3] (C4): synthetic
three-planet near-resonant system, 1500-day baseline, dt = 0.1 d, 10^5 walkers per GPU --
the "3-planet config" the north-star target is quoted on. Weak scaling: every rank
evaluates its own 10^5 walkers; with N > 1 the per-step all-gather of log-likelihoods (the
sampler's exchange step) is inside the timed region.

Prints ONE JSON line (rank 0). `--impl reference` times the CPU implementation of the same
path (the oracle port of the TTVFast path, all host cores) on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SEED = 20240607
# Examples/inputs/trues:13-23 (3-planet true solution; nauyaca order per planet:
# mass[Mearth] period ecc inclination argument mean_anomaly ascending_node)
TRUES_3PL = [0.07652, 10.4707, 0.0406, 89.809, 292.62, 246.14, 88.3,
             1.98033, 14.10509, 0.0244, 89.909, 239.76, 191.31, 88.46,
             0.6705, 23.57467, 0.0306, 90.336, 82.64, 149.24, 88.51]
# Documentation/inputs/true_solution (2-planet system)
TRUES_2PL = [5.66468, 5.35468, 0.1237, 90.7169039033241, 296.22, 16.39, 88.36,
             26.85598, 11.83319, 0.0727, 90.05186862225226, 187.26, 86.62, 88.95]

# Algorithmic FLOP model (+,-,*,/,sqrt = 1, FMA = 2), counted on the canonical arithmetic the
# kernel and the oracle both execute (DESIGN.md "FLOP table"; SURVEY.md 8(d) gave first
# estimates of 200 / 25N(N+1)/2+40N / 6, these are the refined constants):
#   F_KEP  = 185  one universal-variable Kepler drift with the radius handed over by the kick
#                 (setup 39, Stumpff series 33, quintic step 42, Taylor + f,g update 71);
#                 +7 (dot3, sqrt, 1/r) when it computes the radius itself
#   F_R3   = 12   one x/|x|^3  (dot3 5, sqrt 1, 1/r 1, cube 2, scale 3)
#   F_KICK = 9N + 12(2N-1+P) + 15P + 6N + 6 + 21(N-1) + 12N,  P = N(N-1)/2 pairs
#   F_CHK  = 3    per planet and step (sky-plane r.v)
#   F_TR   = 2*((N+1)/2)*192 (synchronise two nodes, mean planet index) + 2*3*211 (two-sided
#            Newton, 3 iterations of drift + derivative) + 35 (weights, bookkeeping)
F_KEP, F_KEP_OWN_R, F_R3, F_CHK = 185.0, 192.0, 12.0, 3.0


def flops_per_eval(npla, nsteps, nconj):
    P = npla * (npla - 1) / 2.0
    f_kick = 9 * npla + F_R3 * (2 * npla - 1 + P) + 15 * P + 6 * npla + 6 + 21 * (npla - 1) + 12 * npla
    f_tr = 2.0 * ((npla + 1) / 2.0) * F_KEP_OWN_R + 2.0 * 3.0 * (F_KEP_OWN_R + 19.0) + 35.0
    return nsteps * (npla * F_KEP + f_kick + npla * F_CHK) + nconj * f_tr


def load_systems():
    with open(os.path.join(ROOT, "tests", "golden", "systems.json")) as fh:
        return json.load(fh)


def cube_of(spec, flat):
    """Hypercube coordinates of a full physical vector (inverse of utils.py:403-438 for the
    free parameters; constants are dropped)."""
    flat = np.asarray(flat, dtype=np.float64)
    const = {int(k) for k in spec["constant_params"]}
    vals = []
    for k in range(7 * spec["npla"]):
        if k in const:
            continue
        c = k % 7
        p = k // 7
        if c == 4:
            v = flat[7 * p + 4] + flat[7 * p + 5]
        elif c == 5:
            v = flat[7 * p + 4] - flat[7 * p + 5]
        else:
            v = flat[k]
        vals.append(v)
    bi, bf = np.array(spec["bi"]), np.array(spec["bf"])
    return (np.array(vals) - bi) / (bf - bi)


def make_workload(name, engine_factory):
    """Returns (spec, X[B, ndim], description). The observed table of the synthetic 1500-d
    system is the engine's own ephemeris of the true solution (computed once, untimed)."""
    systems = load_systems()
    rng = np.random.default_rng(SEED)
    if name == "c2":
        spec = dict(systems["doc2"])
        X = rng.random((10_000, spec["ndim"]))
        return spec, X, "C2: Documentation 2-planet system (221 d, dt 0.18, 1228 steps, 60 obs), 10^4 U(0,1) DE members"
    if name == "c3":
        spec = dict(systems["doc2"])
        c = cube_of(spec, TRUES_2PL)
        X = np.clip(c[None, :] + 0.01 * rng.standard_normal((5000, spec["ndim"])), 0.0, 1.0)
        return spec, X, "C3: 2-planet ptemcee half-step block, 10 temps x 500 walkers, N(x_true,0.01)"
    if name == "c5":
        spec = dict(systems["sys3"])
        X = rng.random((1_000_000, spec["ndim"]))
        return spec, X, "C5 chunk: 3-planet system (500 d, dt 0.34874, 1434 steps, 105 obs), 10^6 U(0,1) walkers of the 10^7 sweep"
    if name != "c4":
        raise SystemExit(f"unknown workload {name}")
    base = dict(systems["sys3f"])
    base["ftime"], base["dt"] = 1500.0, 0.1
    base["observed"] = {}
    base["second_term_sum"] = 0.0
    # pass 1: ephemeris of the true solution over 1500 d on the GPU (untimed set-up)
    eng = engine_factory(base)
    n, pl, ep, tm, rs, vs, st = eng.ephemeris(np.array([TRUES_3PL]), mode=2)
    eng.close()
    k = int(n[0])
    pl, ep, tm, rs = pl[0, :k], ep[0, :k], tm[0, :k], rs[0, :k]
    sig_file = {0: 0.001449, 1: 0.000461, 2: 0.000494}    # Examples/inputs/3pl_planet{1,2,3}_ttvs.dat
    obs, second = {}, 0.0
    for pid, idx in base["planets_IDs"].items():
        m = (pl == idx) & (rs <= base["rstarAU"])
        sg = float(np.sqrt(2.0) * sig_file[idx])           # planetarysystem.py:325 with lo == hi
        obs[pid] = {"planet_index": idx, "epochs": [int(e) for e in ep[m]], "times": [float(t) for t in tm[m]],
                    "sigma": [sg] * int(m.sum())}
        second += float(np.sum(0.5 * np.log(2 * np.pi * np.full(int(m.sum()), sg) ** 2)))
    spec = dict(base)
    spec["observed"] = obs
    spec["second_term_sum"] = second
    c = cube_of(spec, TRUES_3PL)
    X = np.clip(c[None, :] + 0.01 * rng.standard_normal((100_000, spec["ndim"])), 0.0, 1.0)
    nobs = sum(len(o["epochs"]) for o in obs.values())
    return spec, X, (f"C4: synthetic 3-planet near-resonant system, 1500 d, dt 0.1 (15001 steps), {nobs} obs, "
                     "10^5 walkers N(x_true,0.01) per GPU")


# ----------------------------------------------------------------------------- CPU side
def _cpu_worker(args):
    spec, X = args
    from oracle import oracle_py
    osys = oracle_py.OracleSystem(spec, native=True)
    t = time.perf_counter()
    chi2, logl, st = osys.loglike(X, mode=0)
    return time.perf_counter() - t, float(np.sum(logl[np.isfinite(logl)]))


def cpu_rate(spec, X, budget_s, cores):
    """Oracle port over `cores` processes on a bounded sample; returns (evals/s, n_sample). Timed
    on the oracle's NATIVE build (the host's own divide and square root): the checker build emulates
    the GPU's reciprocal seeds by table look-up, which would cost a CPU four times the arithmetic."""
    import multiprocessing as mp
    from oracle import oracle_py
    oracle_py.build()
    osys = oracle_py.OracleSystem(spec, native=True)
    t = time.perf_counter()
    osys.loglike(X[:4], mode=0)
    per = max((time.perf_counter() - t) / 4, 1e-5)
    n = int(max(cores, min(X.shape[0], cores * budget_s / per)))
    n = (n // cores) * cores
    chunks = np.array_split(X[:n], cores)
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_worker, [(spec, c[:2]) for c in chunks])        # warm the workers
        t = time.perf_counter()
        pool.map(_cpu_worker, [(spec, c) for c in chunks])
        wall = time.perf_counter() - t
    return n / wall, n


_PLUMB = {}


def _plumb_init(base, spec):
    import contextlib
    import io
    os.environ["TTVO_NATIVE"] = "1"
    from oracle import ref_plumbing, ref_systems
    with contextlib.redirect_stdout(io.StringIO()):
        nau, U = ref_plumbing.import_reference()
        PS = ref_systems.with_spec(ref_systems.build(nau, ref_plumbing.inputs_dir())[base], spec)
    _PLUMB["f"], _PLUMB["PS"] = U.log_likelihood_func, PS


def _plumb_work(x):
    return _PLUMB["f"](x, _PLUMB["PS"])


def plumbing_rate(workload, spec, X, budget_s, cores):
    """the UNMODIFIED reference Python path (nauyaca.utils.log_likelihood_func,
    utils.py:333-399, from baseline/_ref) under multiprocessing.Pool(cores).map, as mcmc.py:193
    runs it, with the oracle port (native build) standing in for the ttvfast C extension, which
    cannot be installed here. Returns None where the reference package is absent."""
    import multiprocessing as mp
    from oracle import ref_plumbing
    if ref_plumbing.reference_dir() is None:
        return None
    base = {"c2": "doc2", "c3": "doc2", "c5": "sys3", "c4": "sys3f"}[workload]
    ctx = mp.get_context("fork")
    with ctx.Pool(cores, initializer=_plumb_init, initargs=(base, spec)) as pool:
        pool.map(_plumb_work, list(X[:2 * cores]))                            # warm-up (imports, first calls)
        t = time.perf_counter()
        pool.map(_plumb_work, list(X[:2 * cores]))                            # rate estimate
        per = max((time.perf_counter() - t) / 2, 1e-4)                        # seconds per call per core
        n = int(max(cores, min(X.shape[0], cores * budget_s / per)))
        t = time.perf_counter()
        vals = pool.map(_plumb_work, list(X[:n]))
        wall = time.perf_counter() - t
    return {"value": n / wall, "unit": "evals/s", "cores": cores, "kind": "reference-plumbing",
            "sample": f"first {n} walkers, nauyaca.utils.log_likelihood_func (unmodified, baseline/_ref) under "
                      f"multiprocessing.Pool({cores}).map over an oracle-backed ttvfast module",
            "finite": int(np.isfinite(vals).sum())}


def run_reference(args):
    """--impl reference: the CPU implementation of the path on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    systems = load_systems()
    if args.workload == "c4":
        # needs the synthetic observed table: built with the oracle itself here (CPU arm)
        from oracle import oracle_py
        oracle_py.build()
        spec, X, desc = make_workload("c4", lambda s: _OracleEngineShim(s))
    else:
        spec, X, desc = make_workload(args.workload, None)
    cores = os.cpu_count() or 1
    budget = 12.0
    times = []
    n = 0
    for i in range(args.warmup + args.steps):
        rate, n = cpu_rate(spec, X, budget / max(1, args.steps), cores)
        if i >= args.warmup:
            times.append(rate)
    value = float(np.mean(times))
    out = {"impl": "reference", "metric": "likelihood evals/sec", "value": value, "unit": "evals/s",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": 1e3 * n / value, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic",
           "config": {"workload": desc},
           "cpu_baseline": {"value": value, "unit": "evals/s", "cores": cores, "kind": "port",
                            "sample": f"{n} walkers of the workload per step, oracle C port (native divide/sqrt build) in {cores} processes"},
           "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if not args.no_plumbing:
        pl = plumbing_rate(args.workload, spec, X, 8.0, cores)
        if pl is not None:
            out["cpu_baseline"]["plumbing"] = pl
    print(json.dumps(out))


class _OracleEngineShim:
    """CPU arm only: gives make_workload('c4') the true-solution ephemeris from the oracle."""

    def __init__(self, spec):
        self.spec = spec

    def ephemeris(self, X, mode=2):
        from oracle import oracle_py
        s = self.spec
        st, pl, ep, tm, rs, vs = oracle_py.ephemeris(X[0], s["mstar"], s["t0"], s["dt"], s["ftime"])
        k = len(tm)
        return (np.array([k]), pl[None, :], ep[None, :], tm[None, :], rs[None, :], vs[None, :], np.array([st]))

    def close(self):
        pass


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock, power and throttle reasons of one GPU DURING the timed region: an NVML polling thread
    (a sample every 10 ms, the first one at once - `nvidia-smi -lms` needs ~0.5 s to come up on an
    8-GPU box and then misses a short region), nvidia-smi as the fall-back where NVML does not load."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.thread = None
        self.samples = []          # (sm_mhz, max_mhz, watts, reason bits)

    def _nvml_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if self.idx < len(ids) and ids[self.idx].isdigit():
                return int(ids[self.idx])
        return self.idx

    def start(self):
        try:
            import threading
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self._nvml_index())
            mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self._stop = threading.Event()

            def poll():
                while True:
                    try:
                        self.samples.append((float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), mx,
                                             nv.nvmlDeviceGetPowerUsage(h) / 1000.0,
                                             int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))))
                    except Exception:
                        pass
                    if self._stop.wait(0.01):
                        return
            self.thread = threading.Thread(target=poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.thread is not None:
            self._stop.set()
            self.thread.join(timeout=2)
            if not self.samples:
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
            bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
            # the samples under load: the region starts and may end at the idle clock
            sm = [x[0] for x in self.samples]
            reasons = sorted(n for n, b in bits.items() if any(x[3] & b for x in self.samples))
            return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(self.samples[0][1]), "reasons": reasons,
                    "power_w_max": float(max(x[2] for x in self.samples)), "samples": len(sm), "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons, pw = [], [], set(), []
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for nme, v in zip(self.NAMES, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "power_w_max": float(max(pw)), "samples": len(sm), "source": "nvidia-smi"}


# ----------------------------------------------------------------------------- GPU side
def run_gpu(args):
    from nauyaca_b200 import _lib as _ttvb_lib
    _ttvb_lib.ensure_built(builder=int(os.environ.get("LOCAL_RANK", "0")) == 0)
    import torch
    import nauyaca_b200 as nb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (nauyaca_b200 has no CPU fallback)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    spec, X, desc = make_workload(args.workload, lambda s: nb.BatchEngine(s, device=local))
    # weak scaling: every rank draws its own walkers of the same distribution
    if world > 1:
        rng = np.random.default_rng(SEED + rank)
        if args.workload in ("c2", "c5"):
            X = rng.random(X.shape)
        else:
            c = cube_of(spec, TRUES_3PL if spec["npla"] == 3 else TRUES_2PL)
            X = np.clip(c[None, :] + 0.01 * rng.standard_normal(X.shape), 0.0, 1.0)
    B, ndim = X.shape
    eng = nb.BatchEngine(spec, device=local)
    stream = torch.cuda.current_stream().cuda_stream

    Xd = torch.from_numpy(X).cuda()
    chi2 = torch.empty(B, dtype=torch.float64, device="cuda")
    logl = torch.empty(B, dtype=torch.float64, device="cuda")
    stat = torch.empty(B, dtype=torch.int32, device="cuda")
    gathered = torch.empty(world * B, dtype=torch.float64, device="cuda") if world > 1 else None
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    def step_resident():
        flush.zero_()
        eng.loglike_into(Xd.data_ptr(), B, nb.MODE_CUBE, chi2.data_ptr(), logl.data_ptr(), stat.data_ptr(), stream)
        if world > 1:
            dist.all_gather_into_tensor(gathered, logl)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    launches0 = eng.launch_count()
    for _ in range(args.warmup):
        step_resident()
    sync_all()
    clocks = ClockSampler(local)
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms = []
    launches_before = eng.launch_count()
    e0.record()
    for _ in range(args.steps):
        step_resident()
        if args.per_kernel:
            kernel_ms.append(eng.last_kernel_ms())
    e1.record()
    sync_all()
    ms = e0.elapsed_time(e1)
    launches = eng.launch_count() - launches_before
    # kernel-only duration (CUDA events on the launching stream, recorded inside the ABI call)
    if not kernel_ms:
        for _ in range(min(3, args.steps)):
            step_resident()
            kernel_ms.append(eng.last_kernel_ms())
    clk = clocks.stop()
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * B * args.steps / (ms_max * 1e-3)

    # ---- end to end through the public API with HOST buffers (pinned), copies inside
    Xh = torch.from_numpy(X).pin_memory()
    logl_h = torch.empty(B, dtype=torch.float64).pin_memory()
    chi2_h = torch.empty(B, dtype=torch.float64).pin_memory()
    stat_h = torch.empty(B, dtype=torch.int32).pin_memory()

    def step_e2e():
        eng.loglike_into(Xh.data_ptr(), B, nb.MODE_CUBE, chi2_h.data_ptr(), logl_h.data_ptr(), stat_h.data_ptr(), stream)

    for _ in range(max(1, min(args.warmup, 2))):
        step_e2e()
    sync_all()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        step_e2e()
    e1.record()
    sync_all()
    e2e_ms = max(e0.elapsed_time(e1), 0.0)
    t = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * B * args.steps / (float(t.item()) * 1e-3)

    # ---- parity spot-check of what was just timed (host result vs device result, bits)
    same = bool(np.array_equal(logl_h.numpy(), logl.cpu().numpy(), equal_nan=True))

    if rank == 0:
        # mean conjunctions per system for the FLOP model, from a 256-system table call
        n_ev, *_ = eng.ephemeris(eng.cube_to_physical(X[:256])[0], mode=2, cap=1024)
        nconj = float(np.mean(n_ev))
        f_eval = flops_per_eval(spec["npla"], eng.nsteps, nconj)
        kms = float(np.mean(kernel_ms))
        achieved = f_eval * B / (kms * 1e-3) / 1e12
        peak = eng.fp64_peak_tflops()
        nominal = 148 * 64 * 2 * 1.965e9 / 1e12
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
                traffic = json.load(fh).get(args.workload)
        except Exception:
            traffic = None
        out = {
            "metric": "likelihood evals/sec", "value": value, "unit": "evals/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": desc, "batch_per_gpu": B, "ndim": ndim, "npla": spec["npla"],
                       "nsteps": eng.nsteps, "mean_conjunctions": nconj,
                       "l2": "flushed between steps (256 MiB memset inside the timed region)",
                       "collective": "all_gather(logl) per step" if world > 1 else "none"},
            "e2e": {"value": e2e_value, "unit": "evals/s", "h2d_bytes_per_step": int(B * ndim * 8),
                    "d2h_bytes_per_step": int(B * (8 + 8 + 4))},
            "gpu_launches": int(launches),
            "clocks": clk,
            "roofline": {"bound": "fp64", "bound_note": "FP64 vector pipe: this path has no dense contraction and moves ~0.2 GB/s of HBM (north_star: % of FP64 peak), so neither hbm nor tensor applies", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved / peak if peak else None,
                         "traffic": traffic["dram_bytes_per_launch"] if traffic else None,
                         "traffic_source": traffic["source"] if traffic else None,
                         "traffic_note": "above the algorithmic bytes by the context hand-over of time-segmented scheduling (DESIGN.md)",
                         "peak_source": "FP64 FMA-chain microbenchmark in this run (MEASURED_PEAKS.json has no FP64 entry)",
                         "peak_note": "the peak chains fma(c, const, const); with register operands taken from other in-flight "
                                      "chains, as the step loop's DFMAs have them, the same pipe sustains 66-77 % of nominal "
                                      "(profiles/r2y_fp64_register_operands.txt); ncu: pipe 73 % active over a C4 launch "
                                      "(profiles/r2z_c4_full.txt)",
                         "nominal_peak": nominal, "frac_of_nominal": achieved / nominal,
                         "kernel_ms": kms, "flop_per_eval": f_eval,
                         "hbm": {"algorithmic_bytes_per_launch": int(B * (ndim + 2) * 8 + B * 4),
                                 "achieved_gbs": (B * (ndim + 2) * 8 + B * 4) / (kms * 1e-3) / 1e9}},
            "host_device_bits_equal": same,
        }
        if world == 1 and not args.no_cpu:
            cores = os.cpu_count() or 1
            rate, n = cpu_rate(spec, X, 12.0, cores)
            out["cpu_baseline"] = {"value": rate, "unit": "evals/s", "cores": cores, "kind": "port",
                                   "sample": f"first {n} walkers of the same batch, oracle C port (native divide/sqrt build) in {cores} processes"}
            if not args.no_plumbing:
                pl = plumbing_rate(args.workload, spec, X, 8.0, cores)
                if pl is not None:
                    out["cpu_baseline"]["plumbing"] = pl
        print(json.dumps(out))
    eng.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_c3(args):
    from nauyaca_b200 import _lib as _ttvb_lib
    _ttvb_lib.ensure_built(builder=int(os.environ.get("LOCAL_RANK", "0")) == 0)
    """BASELINE configs[2]: the ptemcee-style MCMC itself. A step = one full PT iteration of
    10 temperatures x 1000 walkers (two half-ensemble stretch moves of 5000 proposals each,
    temperature swaps, ladder adaptation) through nauyaca_b200.mcmc.make_sampler; host sampler
    work and the host<->device copies of every half-step are inside the timed region, so
    value == e2e. With N > 1 every rank runs an independent replica of the ensemble (a
    5000-proposal block is latency-bound on one GPU: sharding it does not shorten the step)."""
    import torch
    import nauyaca_b200 as nb
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nb.utils.DEFAULT_DEVICE = local
    spec = dict(load_systems()["doc2"])
    PS = nb.SpecSystem(spec)
    nt, nw = 10, 1000
    shard = bool(args.shard_walkers) and world > 1 and not args.host_sampler
    rng = np.random.mtrand.RandomState(SEED + (0 if shard else rank))
    c = cube_of(spec, TRUES_2PL)
    p0 = np.clip(c[None, None, :] + 0.01 * rng.normal(size=(nt, nw, spec["ndim"])), 0.0, 1.0)
    eng = nb.utils.engine_for(PS, device=local)
    if args.host_sampler:
        sampler = nb.mcmc.make_sampler(PS, p0, tmax=100.0, random=rng)
        it = sampler.sample(p0, iterations=args.warmup + args.steps, thin=1, storechain=False, adapt=True)
        for _ in range(args.warmup):
            next(it)
    else:
        sampler = nb.mcmc.make_device_sampler(PS, p0, tmax=100.0, seed=SEED + (0 if shard else rank), device=local,
                                              shard=shard)
        next(sampler.sample(p0, iterations=1, adapt=True, storechain=False))
        if args.warmup > 1:
            sampler.advance(args.warmup - 1, adapt=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = ClockSampler(local); clocks.start()
    l0 = eng.launch_count()
    t0 = time.perf_counter()
    if args.host_sampler:
        for _ in range(args.steps):
            p, logpost, logl = next(it)
    else:
        p, logpost, logl = sampler.advance(args.steps, adapt=True)      # returns after the state is back on the host
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3
    launches = eng.launch_count() - l0
    kms = eng.last_kernel_ms()
    clk = clocks.stop()
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    evals = (1 if shard else world) * args.steps * nt * nw
    if rank == 0:
        B = nt * nw // 2
        f_eval = flops_per_eval(spec["npla"], eng.nsteps, 60.0)
        value = evals / (ms * 1e-3)
        peak = eng.fp64_peak_tflops()
        achieved = f_eval * B / (kms * 1e-3) / 1e12
        out = {"metric": "likelihood evals/sec", "value": value, "unit": "evals/s", "n_gpus": world,
               "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
               "scaling": "strong" if shard else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "config": {"workload": "C3: ptemcee-style PT MCMC, 2-planet system (1228 steps, 60 obs), 10 temperatures x 1000 "
                                      "walkers, one PT iteration per step (2 x 5000 proposals), "
                                      + ("ONE ensemble, proposals sharded over the GPUs, all_gather(logl) per half-step"
                                         if shard else "independent replica per GPU") + "; sampler moves "
                                      + ("on the host (numpy)" if args.host_sampler else "on the device (ttvb_pt_*)"),
                          "mean_logl_cold": float(np.mean(logl[0])), "swap_fraction_cold": float(sampler.tswap_acceptance_fraction[0]),
                          "l2": "not flushed: a 5000-proposal block is 0.5 MB and latency-bound"},
               "e2e": {"value": value, "unit": "evals/s",
                       "h2d_bytes_per_step": int(2 * B * spec["ndim"] * 8) if args.host_sampler else 0,
                       "d2h_bytes_per_step": int(2 * B * 20) if args.host_sampler
                       else int(nt * nw * (spec["ndim"] + 3) * 8 / args.steps)},
               "gpu_launches": int(launches), "clocks": clk,
               "roofline": {"bound": "fp64", "bound_note": "FP64 vector pipe: this path has no dense contraction and moves ~0.2 GB/s of HBM (north_star: % of FP64 peak), so neither hbm nor tensor applies", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                            "frac": achieved / peak, "traffic": None, "kernel_ms": kms, "flop_per_eval": f_eval,
                            "note": "latency-bound: 157 warps on 592 schedulers"}}
        print(json.dumps(out))
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


def run_c5_strong(args):
    """BASELINE configs[4] as a STRONG-scaling job: the 10^7-point uniform sweep of the 3-planet
    system (init_walkers / parameter-grid shape, utils.py:463-543), fixed total, split into
    contiguous shards over the ranks. A step = the whole sweep, end to end: every rank streams
    its shard from PINNED HOST memory through the C ABI (chunked, uploads and downloads under
    the kernels), log-likelihoods come back to the host, and ONE all-gather at the end gives every
    rank all 10^7 values (the sweep's only exchange). value == e2e by construction."""
    from nauyaca_b200 import _lib as _ttvb_lib
    _ttvb_lib.ensure_built(builder=int(os.environ.get("LOCAL_RANK", "0")) == 0)
    import torch
    import nauyaca_b200 as nb
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    spec = dict(load_systems()["sys3"])
    total = int(args.sweep_points)
    lo, hi = nb.shard_bounds(total, world, rank)
    B, ndim = hi - lo, spec["ndim"]
    # row i of the sweep is the same whatever the number of ranks: blocks of 10^6 rows, one seed each
    BLK = 1_000_000
    Xh = torch.empty((B, ndim), dtype=torch.float64).pin_memory()
    Xn = Xh.numpy()
    for blk in range(lo // BLK, (hi - 1) // BLK + 1):
        rows = np.random.default_rng([SEED, blk]).random((min(BLK, total - blk * BLK), ndim))
        a, b = max(lo, blk * BLK), min(hi, (blk + 1) * BLK)
        Xn[a - lo:b - lo] = rows[a - blk * BLK:b - blk * BLK]
    logl_h = torch.empty(B, dtype=torch.float64).pin_memory()
    stat_h = torch.empty(B, dtype=torch.int32).pin_memory()
    eng = nb.BatchEngine(spec, device=local)
    stream = torch.cuda.current_stream().cuda_stream
    maxb = nb.shard_bounds(total, world, 0)[1]
    mine = torch.zeros(maxb, dtype=torch.float64, device="cuda")
    gathered = torch.empty(world * maxb, dtype=torch.float64, device="cuda") if world > 1 else None

    def sweep():
        eng.loglike_into(Xh.data_ptr(), B, nb.MODE_CUBE, None, logl_h.data_ptr(), stat_h.data_ptr(), stream)
        if world > 1:
            mine[:B].copy_(logl_h, non_blocking=True)
            dist.all_gather_into_tensor(gathered, mine)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        sweep()
    sync_all()
    clocks = ClockSampler(local); clocks.start()
    l0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        sweep()
    e1.record()
    sync_all()
    ms = e0.elapsed_time(e1)
    launches = eng.launch_count() - l0
    clk = clocks.stop()
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    # what the kernels alone take on this rank's shard (resident input), for the roofline line
    Xd = torch.from_numpy(Xn[:min(B, 1_000_000)]).cuda()
    o1 = torch.empty(Xd.shape[0], dtype=torch.float64, device="cuda"); o2 = torch.empty_like(o1)
    o3 = torch.empty(Xd.shape[0], dtype=torch.int32, device="cuda")
    eng.loglike_into(Xd.data_ptr(), Xd.shape[0], nb.MODE_CUBE, o1.data_ptr(), o2.data_ptr(), o3.data_ptr(), stream)
    torch.cuda.synchronize()
    kms = eng.last_kernel_ms()
    if rank == 0:
        value = total * args.steps / (ms * 1e-3)
        n_ev, *_ = eng.ephemeris(eng.cube_to_physical(Xn[:256])[0], mode=2, cap=1024)
        f_eval = flops_per_eval(spec["npla"], eng.nsteps, float(np.mean(n_ev)))
        peak = eng.fp64_peak_tflops()
        achieved = f_eval * Xd.shape[0] / (kms * 1e-3) / 1e12
        out = {"metric": "likelihood evals/sec", "value": value, "unit": "evals/s", "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
               "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "config": {"workload": f"C5 sweep: 3-planet system (500 d, dt 0.34874, 1434 steps, 105 obs), {total} U(0,1) points "
                                      f"in total, contiguous shards over {world} GPU(s), host buffers in and out, one "
                                      "all_gather(logl) at the end of the sweep",
                          "points_total": total, "points_per_gpu": B, "ndim": ndim, "npla": spec["npla"], "nsteps": eng.nsteps,
                          "l2": "inputs (1.44 GB in total) far larger than L2",
                          "collective": "all_gather(logl) once per sweep" if world > 1 else "none"},
               "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": int(B * ndim * 8),
                       "d2h_bytes_per_step": int(B * 12)},
               "gpu_launches": int(launches), "clocks": clk,
               "roofline": {"bound": "fp64", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                            "traffic": None, "kernel_ms": kms, "kernel_batch": int(Xd.shape[0]), "flop_per_eval": f_eval,
                            "note": "kernel alone on a resident chunk of this rank's shard"}}
        print(json.dumps(out))
    eng.close()
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c4", choices=["c2", "c3", "c4", "c5"])
    ap.add_argument("--per-kernel", action="store_true", help="read the kernel time after every step (adds a sync)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-plumbing", action="store_true",
                    help="skip the reference-plumbing CPU leg (unmodified nauyaca.utils under multiprocessing)")
    ap.add_argument("--host-sampler", action="store_true", help="c3: sampler moves in numpy on the host instead of on the device")
    ap.add_argument("--shard-walkers", action="store_true",
                    help="c3 with N > 1: ONE ensemble, the proposals of every half-step sharded over the GPUs and their "
                         "log-likelihoods all-gathered over NCCL (strong scaling) instead of one replica per GPU")
    ap.add_argument("--strong", action="store_true",
                    help="c5: the fixed 10^7-point sweep split over the GPUs (strong scaling, host buffers, one gather)")
    ap.add_argument("--sweep-points", type=int, default=10_000_000, help="c5 --strong: total points of the sweep")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "c5" and args.strong:
        run_c5_strong(args)
    elif args.workload == "c3":
        run_c3(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
