"""GPU: planet counts other than the goldens', hostile inputs, status bits, chunked batches.
Always GPU (through the C ABI) against the oracle, bit for bit."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _spec(npla, ftime=60.0, dt=0.05, mstar=1.0, nobs_per=6, seed=0):
    """Synthetic system: all 7*npla parameters free, wide bounds, a small observed table."""
    rng = np.random.default_rng(seed)
    ndim = 7 * npla
    lo = np.tile([0.1, 2.0, 0.0, 80.0, 0.0, 0.0, 0.0], npla).astype(float)
    hi = np.tile([300.0, 40.0, 0.6, 100.0, 360.0, 360.0, 360.0], npla).astype(float)
    for i in range(npla):
        lo[7 * i + 1], hi[7 * i + 1] = 3.0 * 1.7 ** i, 4.0 * 1.7 ** i       # separated periods
    bi, bf = lo.copy(), hi.copy()
    for i in range(npla):                                                   # (w+M, w-M) parameterisation
        bi[7 * i + 4], bf[7 * i + 4] = 0.0, 720.0
        bi[7 * i + 5], bf[7 * i + 5] = -360.0, 360.0
    obs = {}
    for i in range(npla):
        P = 3.5 * 1.7 ** i
        ep = list(range(nobs_per))
        obs[f"p{i}"] = {"planet_index": i, "epochs": ep, "times": [float(1.0 + P * e + 0.01 * rng.normal()) for e in ep],
                        "sigma": [0.002] * nobs_per}
    return {"mstar": mstar, "rstarAU": 0.00465 * 1.0, "t0": 0.0, "ftime": ftime, "dt": dt, "npla": npla, "ndim": ndim,
            "planets_IDs": {f"p{i}": i for i in range(npla)},
            "bounds": [[float(a), float(b)] for a, b in zip(lo, hi)], "bi": list(map(float, bi)), "bf": list(map(float, bf)),
            "hypercube": [[0.0, 1.0]] * ndim, "constant_params": {}, "observed": obs, "second_term_sum": 1.25}


def _same(a, b):
    a = np.asarray(a); b = np.asarray(b)
    ok = (a == b) | (np.isnan(a) & np.isnan(b))
    assert ok.all(), f"{(~ok).sum()} of {a.size} differ: {a[~ok][:3]} vs {b[~ok][:3]}"


@pytest.mark.parametrize("npla", [1, 2, 4, 5, 7, 9])
def test_other_planet_counts(oracle, npla):
    import nauyaca_b200 as nb
    spec = _spec(npla)
    eng = nb.BatchEngine(spec); osys = oracle.OracleSystem(spec)
    X = np.random.default_rng(npla).random((200, spec["ndim"]))
    chi2, logl, st = eng.loglike(X, mode=0)
    o = osys.loglike(X, mode=0)
    _same(st, o[2]); _same(chi2, o[0]); _same(logl, o[1])
    flat, _ = eng.cube_to_physical(X[:16])
    n, pl, ep, tm, rs, vs, st2 = eng.ephemeris(flat, mode=2, cap=600)
    for b in range(16):
        e = oracle.ephemeris(flat[b], spec["mstar"], spec["t0"], spec["dt"], spec["ftime"], cap=600)
        assert n[b] == len(e[3]) and st2[b] == e[0]
        _same(tm[b, :n[b]], e[3]); _same(pl[b, :n[b]], e[1]); _same(ep[b, :n[b]], e[2])
    eng.close()


def test_hostile_elements_and_status_bits(oracle):
    """Huge masses, e up to 0.95, coarse steps: crossing orbits, ejections, Kepler failures.
    Whatever happens must happen identically in the oracle; nothing aborts the batch."""
    import nauyaca_b200 as nb
    spec = _spec(3, ftime=120.0, dt=0.4)
    eng = nb.BatchEngine(spec); osys = oracle.OracleSystem(spec)
    rng = np.random.default_rng(99)
    B = 600
    flat = np.empty((B, 21))
    for i in range(3):
        flat[:, 7 * i + 0] = 10 ** rng.uniform(-1, 4.2, B)          # up to ~16000 Mearth
        flat[:, 7 * i + 1] = rng.uniform(1.5, 9.0, B)               # overlapping periods
        flat[:, 7 * i + 2] = rng.uniform(0.0, 0.95, B)
        flat[:, 7 * i + 3] = rng.uniform(60, 120, B)
        flat[:, 7 * i + 4:7 * i + 7] = rng.uniform(-400, 800, (B, 3))
    flat[0, 2] = 1.5                 # invalid eccentricity
    flat[1, 1] = 0.0                 # invalid period (a = 0)
    flat[2, 0] = np.nan
    chi2, logl, st = eng.loglike(flat, mode=2)
    o = osys.loglike(flat, mode=2)
    _same(st, o[2]); _same(chi2, o[0]); _same(logl, o[1])
    assert st[0] & 16 and st[1] & 16 and st[2] & 16
    assert np.all(np.isfinite(chi2[:2]))                             # penalties, not NaN
    assert (st & 2).any() and (st & 4).any()                         # the set really is hostile
    eng.close()


def test_event_overflow_bit(oracle):
    import nauyaca_b200 as nb
    spec = _spec(2)
    eng = nb.BatchEngine(spec)
    X = np.random.default_rng(5).random((8, spec["ndim"]))
    flat, _ = eng.cube_to_physical(X)
    n, pl, ep, tm, rs, vs, st = eng.ephemeris(flat, mode=2, cap=5)
    assert np.all(n == 5) and np.all(st & 8)
    for b in range(8):
        e = oracle.ephemeris(flat[b], spec["mstar"], spec["t0"], spec["dt"], spec["ftime"], cap=5)
        assert e[0] & 8
        _same(tm[b, :5], e[3])
    eng.close()


def test_identical_walkers_fill_the_queue(oracle):
    """All 32 lanes of a warp trigger in the same step (identical systems): the event queue
    takes 32 pushes per planet per step."""
    import nauyaca_b200 as nb
    spec = _spec(3)
    eng = nb.BatchEngine(spec); osys = oracle.OracleSystem(spec)
    x = np.random.default_rng(1).random(spec["ndim"])
    X = np.tile(x, (300, 1))
    chi2, logl, st = eng.loglike(X, mode=0)
    o = osys.loglike(x[None, :], mode=0)
    assert np.all(chi2 == o[0][0]) and np.all(st == o[2][0])
    eng.close()


def test_queue_fills_in_the_middle_of_a_push(oracle):
    """Warps of 27-31 identical systems plus a few different ones: the 64-slot queue fills while
    a round of simultaneous triggers is being pushed, so the step loop hands the batch over and
    resumes the push (and the last batch of every warp is a partial one with an odd count).
    chi2, status and the event table stay the oracle's, bit for bit."""
    import nauyaca_b200 as nb
    spec = _spec(3, ftime=80.0)
    eng = nb.BatchEngine(spec); osys = oracle.OracleSystem(spec)
    rng = np.random.default_rng(7)
    base = rng.random((4, spec["ndim"]))
    rows = []
    for w in range(8):                       # 8 warps: 32 - n_other copies of one system, n_other others
        n_other = 1 + w % 5
        rows += [base[w % 4]] * (32 - n_other) + [rng.random(spec["ndim"]) for _ in range(n_other)]
    X = np.array(rows)
    chi2, logl, st = eng.loglike(X, mode=0)
    o = osys.loglike(X, mode=0)
    _same(chi2, o[0]); _same(logl, o[1]); _same(st, o[2])
    flat, inside = eng.cube_to_physical(X)
    cap = 256
    n, pl, ep, tm, rs, vs, st2 = eng.ephemeris(flat, mode=2, cap=cap)
    for b in range(0, X.shape[0], 3):
        ost, opl, oep, otm, ors, ovs = oracle.ephemeris(flat[b], spec["mstar"], spec["t0"], spec["dt"], spec["ftime"],
                                                        cap=cap)
        k = n[b]
        assert st2[b] == ost and k == len(otm)
        _same(pl[b, :k], opl); _same(ep[b, :k], oep); _same(tm[b, :k], otm); _same(rs[b, :k], ors); _same(vs[b, :k], ovs)
    eng.close()


def test_two_planets_trigger_in_the_same_step(oracle):
    """Systems whose two inner planets transit in the same step (same period and phase, orbits
    tilted apart), 31 copies per warp plus one unrelated system: a lane pushes two events at
    once, and the queue runs out of room in front of such a pair (odd fill level). Bit for bit
    against the oracle, chi2 and event table."""
    import nauyaca_b200 as nb
    spec = _spec(3, ftime=70.0)
    eng = nb.BatchEngine(spec); osys = oracle.OracleSystem(spec)
    twin = np.array([0.05, 5.0, 0.02, 88.0, 30.0, 100.0, 10.0,
                     0.05, 5.0, 0.02, 92.0, 30.0, 100.0, 60.0,
                     3.0, 11.3, 0.05, 89.5, 200.0, 40.0, 0.0])
    rng = np.random.default_rng(3)
    other = np.array([[2.0, 3.1 + 0.3 * k, 0.03, 89.8, 10.0 * k, 50.0 + 20 * k, 0.0,
                       4.0, 6.9 + 0.2 * k, 0.04, 90.2, 80.0, 120.0 + 7 * k, 0.0,
                       6.0, 13.0 + 0.5 * k, 0.01, 90.0, 150.0, 10.0 * k, 0.0] for k in range(4)])
    rows = []
    for w in range(4):
        rows += [twin] * 31 + [other[w]]
    X = np.array(rows)
    chi2, logl, st = eng.loglike(X, mode=2)
    o = osys.loglike(X, mode=2)
    _same(chi2, o[0]); _same(logl, o[1]); _same(st, o[2])
    n, pl, ep, tm, rs, vs, st2 = eng.ephemeris(X, mode=2, cap=128)
    for b in (0, 30, 31, 32, 127):
        ost, opl, oep, otm, ors, ovs = oracle.ephemeris(X[b], spec["mstar"], spec["t0"], spec["dt"], spec["ftime"], cap=128)
        k = n[b]
        assert st2[b] == ost and k == len(otm)
        _same(pl[b, :k], opl); _same(ep[b, :k], oep); _same(tm[b, :k], otm); _same(rs[b, :k], ors)
    # the twins really do transit in the same step: equal times for planets 0 and 1
    k = n[0]
    t0 = tm[0, :k][pl[0, :k] == 0]; t1 = tm[0, :k][pl[0, :k] == 1]
    assert len(t0) > 5 and len(t0) == len(t1) and np.max(np.abs(t0 - t1)) < 0.02
    eng.close()


def test_chunked_host_batches(oracle):
    """More rows than the 2^20 staging chunk: host-pointer path loops over chunks."""
    import nauyaca_b200 as nb
    spec = _spec(1, ftime=3.0, dt=0.25, nobs_per=1)
    eng = nb.BatchEngine(spec); osys = oracle.OracleSystem(spec)
    B = (1 << 20) + 4097
    X = np.random.default_rng(2).random((B, spec["ndim"]))
    chi2, logl, st = eng.loglike(X, mode=0)
    idx = np.r_[0:50, (1 << 20) - 25:(1 << 20) + 25, B - 50:B]
    o = osys.loglike(X[idx], mode=0)
    _same(chi2[idx], o[0]); _same(st[idx], o[2])
    assert np.isfinite(chi2).all()
    eng.close()


@pytest.mark.parametrize("nseg", [2, 3, 5])
def test_time_segmented_launch_is_bit_identical(oracle, systems, monkeypatch, nseg):
    """A launch cut into time segments (tiles handed from CTA to CTA through the context
    arrays) gives the bits of the unsegmented launch and of the oracle, including events that
    trigger right at a segment boundary."""
    import nauyaca_b200 as nb
    spec = systems["sys3"]; osys = oracle.OracleSystem(spec)
    X = np.random.default_rng(31).random((700, spec["ndim"]))
    X[5, 0] = 2.0
    monkeypatch.setenv("TTVB_NSEG", str(nseg))
    eng = nb.BatchEngine(spec)
    chi2, logl, st, per = eng.loglike(X, mode=0, per_planet=True)
    o = osys.loglike(X, mode=0)
    _same(st, o[2]); _same(chi2, o[0]); _same(logl, o[1])
    flat, _ = eng.cube_to_physical(X[:40])
    n, pl, ep, tm, rs, vs, st2 = eng.ephemeris(flat, mode=2, cap=200)
    for b in range(40):
        e = oracle.ephemeris(flat[b], spec["mstar"], spec["t0"], spec["dt"], spec["ftime"], cap=200)
        assert n[b] == len(e[3])
        _same(tm[b, :n[b]], e[3]); _same(ep[b, :n[b]], e[2])
    eng.close()


@pytest.mark.parametrize("lanes,nseg", [(1, 0), (5, 0), (9, 3), (16, 0), (27, 2), (32, 0), (0, 0)])
def test_thin_launch_is_bit_identical(oracle, systems, monkeypatch, lanes, nseg):
    """A thin launch (systems on the first `lanes` lanes of every warp only: what plan_launch picks for
    batches of at most one warp per scheduler; 0 = its own choice) gives the bits of the full-warp launch
    and of the oracle - chi2, status, per-planet values and the event table, also when cut into time
    segments, with an out-of-bounds row and a ragged last tile."""
    import nauyaca_b200 as nb
    spec = systems["sys3"]; osys = oracle.OracleSystem(spec)
    X = np.random.default_rng(32).random((701, spec["ndim"]))
    X[5, 0] = 2.0
    if lanes:
        monkeypatch.setenv("TTVB_LANES", str(lanes))
    if nseg:
        monkeypatch.setenv("TTVB_NSEG", str(nseg))
    eng = nb.BatchEngine(spec)
    chi2, logl, st, per = eng.loglike(X, mode=0, per_planet=True)
    o = osys.loglike(X, mode=0)
    _same(st, o[2]); _same(chi2, o[0]); _same(logl, o[1])
    import torch
    r = eng.loglike(torch.from_numpy(X[:333]).cuda(), mode=0)            # device pointers, another batch size
    _same(r[0].cpu().numpy(), o[0][:333])
    flat, _ = eng.cube_to_physical(X[:41])
    n, pl, ep, tm, rs, vs, st2 = eng.ephemeris(flat, mode=2, cap=200)
    for b in range(41):
        e = oracle.ephemeris(flat[b], spec["mstar"], spec["t0"], spec["dt"], spec["ftime"], cap=200)
        assert n[b] == len(e[3])
        _same(tm[b, :n[b]], e[3]); _same(ep[b, :n[b]], e[2]); _same(pl[b, :n[b]], e[1])
    eng.close()


def test_launch_between_one_and_two_ctas_per_sm(oracle, systems, monkeypatch):
    """20 000 systems need two CTAs on some SMs: the launch planner spreads them evenly over two CTAs on
    every SM (17 lanes per warp). Same bits as the full-warp launch; a sample against the oracle."""
    import nauyaca_b200 as nb
    spec = systems["doc2"]; osys = oracle.OracleSystem(spec)
    X = np.random.default_rng(33).random((20000, spec["ndim"]))
    eng = nb.BatchEngine(spec)
    chi2, logl, st = eng.loglike(X, mode=0)
    eng.close()
    monkeypatch.setenv("TTVB_LANES", "32")
    eng = nb.BatchEngine(spec)
    c32, l32, s32 = eng.loglike(X, mode=0)
    eng.close()
    _same(chi2, c32); _same(st, s32); _same(logl, l32)
    idx = np.r_[0:150, 9900:10050, 19850:20000]
    o = osys.loglike(X[idx], mode=0)
    _same(chi2[idx], o[0]); _same(st[idx], o[2])


def test_repeated_launches_give_the_same_bits(systems, monkeypatch):
    """A race (inter-CTA hand-over of time segments, event queues, staging / ring aliasing) would show up
    as a result that changes from launch to launch: the same 40 000 systems twelve times, cut into three
    time segments, through the device-pointer and the host-pointer path, must give ONE set of bits
    (scripts/determinism_check.py does the same for the benchmark shapes)."""
    import hashlib
    import torch
    import nauyaca_b200 as nb
    spec = systems["sys3"]
    monkeypatch.setenv("TTVB_NSEG", "3")
    eng = nb.BatchEngine(spec)
    X = np.random.default_rng(34).random((40000, spec["ndim"]))
    Xd = torch.from_numpy(X).cuda()
    seen = set()
    for i in range(12):
        if i % 3 == 2:
            chi2, logl, st = eng.loglike(X, mode=0)
        else:
            c, l, s_ = eng.loglike(Xd, mode=0)
            torch.cuda.synchronize()
            chi2, st = c.cpu().numpy(), s_.cpu().numpy()
        seen.add(hashlib.sha1(chi2.tobytes() + st.tobytes()).hexdigest())
    assert len(seen) == 1
    eng.close()


def test_observed_table_variants(oracle):
    """Planets without observations, unsorted epoch order in the table, epochs the simulation
    never reaches (1e20 each, utils.py:199-206), barycentric-Julian-date sized times."""
    import copy
    import nauyaca_b200 as nb
    base = _spec(3, ftime=50.0, dt=0.06)
    X = np.random.default_rng(8).random((160, base["ndim"]))
    # (a) middle planet has no TTVs
    s = copy.deepcopy(base); del s["observed"]["p1"]
    eng = nb.BatchEngine(s); o = oracle.OracleSystem(s).loglike(X, mode=0)
    r = eng.loglike(X, mode=0, per_planet=True)
    _same(r[0], o[0]); _same(r[2], o[2]); assert np.all(r[3][:, 1] == 0.0); eng.close()
    # (b) epochs far beyond the simulated span: every one costs 1e20
    s = copy.deepcopy(base); s["observed"]["p0"]["epochs"] = [0, 1, 2, 500, 501, 502]
    eng = nb.BatchEngine(s); o = oracle.OracleSystem(s).loglike(X, mode=0)
    r = eng.loglike(X, mode=0); _same(r[0], o[0]); assert np.all(r[0] >= 3e20); eng.close()
    # (c) table not in epoch order: same set of terms; the sum order may differ in the last bit
    s = copy.deepcopy(base)
    for k in ("epochs", "times", "sigma"):
        s["observed"]["p2"][k] = s["observed"]["p2"][k][::-1]
    eng = nb.BatchEngine(s); o = oracle.OracleSystem(s).loglike(X, mode=0)
    r = eng.loglike(X, mode=0)
    assert np.allclose(r[0], o[0], rtol=1e-14, atol=0); eng.close()
    # (d) BJD-sized reference time (miscellany.ipynb cell 61: t0 = 2458690)
    s = copy.deepcopy(base); s["t0"] = 2458690.0; s["ftime"] = 2458740.0
    for p_ in s["observed"].values():
        p_["times"] = [t + 2458690.0 for t in p_["times"]]
    eng = nb.BatchEngine(s); o = oracle.OracleSystem(s).loglike(X, mode=0)
    r = eng.loglike(X, mode=0); _same(r[0], o[0]); _same(r[2], o[2]); eng.close()
    # (e) duplicate epochs are refused at creation
    s = copy.deepcopy(base); s["observed"]["p0"]["epochs"][1] = s["observed"]["p0"]["epochs"][0]
    with pytest.raises(nb.TTVBError):
        nb.BatchEngine(s)


def test_physical_mode_bounds(oracle):
    import nauyaca_b200 as nb
    spec = _spec(2)
    eng = nb.BatchEngine(spec); osys = oracle.OracleSystem(spec)
    lo = np.array([b[0] for b in spec["bounds"]]); hi = np.array([b[1] for b in spec["bounds"]])
    X = lo + np.random.default_rng(3).random((64, spec["ndim"])) * (hi - lo)
    X[5, 2] = hi[2] + 1e-9; X[6, 0] = lo[0] - 1e-9; X[7] = lo; X[8] = hi
    r = eng.loglike(X, mode=1); o = osys.loglike(X, mode=1)
    _same(r[0], o[0]); _same(r[1], o[1]); _same(r[2], o[2])
    assert r[2][5] & 1 and r[2][6] & 1 and not (r[2][7] & 1) and not (r[2][8] & 1)
    eng.close()


def test_two_engines_of_different_table_size_share_a_kernel(systems, oracle):
    """The dynamic shared-memory limit is a property of the kernel function, not of a handle: creating
    a second engine with a smaller observed table (here: none, as utils.run_TTVFast does) must not
    break the launches of the first."""
    import nauyaca_b200 as nb
    spec = systems["sys3f"]
    big = nb.BatchEngine(spec)
    X = np.random.default_rng(3).random((8, spec["ndim"]))
    ref = big.loglike(X, mode=0)[0]
    raw = nb.utils._raw_engine(3, spec["mstar"], spec["t0"], spec["ftime"], spec["dt"])
    raw.ephemeris(oracle.OracleSystem(spec).cube_to_physical(X[0])[None, :], mode=2)
    assert np.array_equal(big.loglike(X, mode=0)[0], ref)
    big.close()


def test_streamed_host_path_and_its_fallback(systems, oracle):
    """Host-pointer calls of two waves or more are streamed (tiles wait on the device for rows the copy
    stream uploads meanwhile). Same bits as the resident path; and when the uploads cannot run under the
    kernel (what a serialising profiler does; forced here with TTVB_STREAM_TEST_STALL) the launch gives up
    after a bounded wait, is run again, and streaming is switched off for the handle."""
    import json
    import os
    import subprocess
    import sys
    code = r'''
import json, sys, numpy as np
sys.path.insert(0, %r)
import torch, nauyaca_b200 as nb
spec = json.load(open(%r))["doc2"]
eng = nb.BatchEngine(spec)
X = np.random.default_rng(5).random((90000, spec["ndim"]))         # 2.4 waves: streamed
res = eng.loglike(torch.from_numpy(X).cuda(), mode=0); torch.cuda.synchronize()
a = eng.loglike(X, mode=0); r1 = eng.stream_retries()
b = eng.loglike(X, mode=0); r2 = eng.stream_retries()
ok = all(np.array_equal(u, v.cpu().numpy()) for u, v in zip(a, res)) and all(np.array_equal(u, v) for u, v in zip(a, b))
print(json.dumps({"ok": bool(ok), "retries": [r1, r2], "status_max": int(a[2].max())}))
''' % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
       os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "systems.json"))
    for stall, want in (("0", [0, 0]), ("1", [1, 1])):
        out = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, TTVB_STREAM_TEST_STALL=stall),
                             capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stderr[-1500:]
        rep = json.loads(out.stdout.strip().splitlines()[-1])
        assert rep["ok"] and rep["retries"] == want and rep["status_max"] == 0, (stall, rep)
