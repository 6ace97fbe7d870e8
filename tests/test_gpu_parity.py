"""GPU: the CUDA path (through the C ABI) against the CPU oracle on the same inputs.

Bar (north star): transit times within 1e-7 d, chi2 within 1e-9 relative, counts/epochs
exact. The kernel and the oracle execute the same IEEE-754 operation sequence, so these
tests assert BIT equality and would also pass the stated tolerances."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL_TIME = 1e-7
TOL_CHI2 = 1e-9


@pytest.fixture(scope="module")
def engines(systems):
    import nauyaca_b200 as nb
    return {k: nb.BatchEngine(v) for k, v in systems.items()}


def _cubes(spec, n, seed, known=None, center=None):
    rng = np.random.default_rng(seed)
    X = rng.random((n, spec["ndim"]))
    return X


def _assert_same(a, b):
    a = np.asarray(a); b = np.asarray(b)
    same = (a == b) | (np.isnan(a) & np.isnan(b))
    assert same.all(), f"{(~same).sum()} of {a.size} differ; first {a[~same][:3]} vs {b[~same][:3]}"


@pytest.mark.parametrize("name,n", [("doc2", 1000), ("sys3", 1000), ("sys3f", 300)])
def test_loglike_uniform_cube_bitwise(engines, oracle, systems, name, n):
    spec = systems[name]; eng = engines[name]; osys = oracle.OracleSystem(spec)
    X = _cubes(spec, n, 20240607)
    chi2, logl, st = eng.loglike(X, mode=0)
    ochi2, ologl, ost = osys.loglike(X, mode=0)
    _assert_same(st, ost); _assert_same(chi2, ochi2); _assert_same(logl, ologl)
    assert np.allclose(chi2, ochi2, rtol=TOL_CHI2, atol=0)


def test_loglike_near_mode_walkers(engines, oracle, systems, known):
    # realistic walkers around the G1 solution (small residuals: the hard case for 1e-9)
    spec = systems["sys3"]; eng = engines["sys3"]; osys = oracle.OracleSystem(spec)
    phys = np.array(known["G1"]["phys_free"])
    bi, bf = np.array(spec["bi"]), np.array(spec["bf"])
    # invert cube_to_physical for the centre: (w+M, w-M) parameterisation of slots 4,5 of each planet
    flat = np.array(known["G1"]["flat_params"])
    names = spec["params_names"].split()
    cube = np.empty(spec["ndim"])
    for i, nm in enumerate(names):
        pl = int(nm[-1]) - 1
        if nm.startswith("argument"):
            v = flat[7 * pl + 4] + flat[7 * pl + 5]
        elif nm.startswith("mean_anomaly"):
            v = flat[7 * pl + 4] - flat[7 * pl + 5]
        else:
            v = phys[i]
        cube[i] = (v - bi[i]) / (bf[i] - bi[i])
    rng = np.random.default_rng(5)
    X = np.clip(cube[None, :] + 1e-4 * rng.standard_normal((512, spec["ndim"])), 0, 1)
    X[0] = cube
    chi2, logl, st = eng.loglike(X, mode=0)
    ochi2, ologl, ost = osys.loglike(X, mode=0)
    _assert_same(st, ost); _assert_same(chi2, ochi2); _assert_same(logl, ologl)
    assert chi2[0] < 1e4          # the centre really is a good fit


@pytest.mark.parametrize("g", ["G1", "G2"])
def test_golden_ephemerides_on_gpu(engines, systems, known, g):
    k = known[g]; spec = systems[k["system"]]; eng = engines[k["system"]]
    n, pl, ep, tm, rs, vs, st = eng.ephemeris(np.array([k["flat_params"]]), mode=2)
    assert st[0] == 0 and n[0] == 105
    assert tm[0, n[0]] == -2.0                       # TTVFast DEFAULT padding (utils.py:76-79)
    pl, ep, tm, rs = pl[0, :n[0]], ep[0, :n[0]], tm[0, :n[0]], rs[0, :n[0]]
    for pid, idx in spec["planets_IDs"].items():
        m = (pl == idx) & (rs <= spec["rstarAU"])
        ge = k["ephemeris"][pid]
        assert list(ep[m]) == ge["epochs"]
        if ge["times"]:
            assert np.max(np.abs(tm[m] - np.array(ge["times"]))) <= TOL_TIME


@pytest.mark.parametrize("g,tol", [("G1", TOL_CHI2), ("G2", 0.0), ("G3", TOL_CHI2), ("G4", 1e-6)])
def test_golden_chi2_on_gpu(engines, known, g, tol):
    k = known[g]; eng = engines[k["system"]]
    chi2, logl, st, per = eng.loglike(np.array([k["flat_params"]]), mode=2, per_planet=True)
    assert abs(chi2[0] - k["chi2_current"]) <= tol * k["chi2_current"]
    assert chi2[0] == per[0].sum() or abs(chi2[0] - per[0].sum()) < 1e-9 * chi2[0]


@pytest.mark.parametrize("name", ["doc2", "sys3"])
def test_ephemeris_table_bitwise(engines, oracle, systems, name):
    spec = systems[name]; eng = engines[name]; osys = oracle.OracleSystem(spec)
    X = _cubes(spec, 96, 3)
    flat, inside = eng.cube_to_physical(X)
    n, pl, ep, tm, rs, vs, st = eng.ephemeris(flat, mode=2, cap=400)
    for b in range(X.shape[0]):
        assert list(flat[b]) == list(osys.cube_to_physical(X[b]))
        o = oracle.ephemeris(flat[b], spec["mstar"], spec["t0"], spec["dt"], spec["ftime"], cap=400)
        assert st[b] == o[0] and n[b] == len(o[3])
        k = n[b]
        _assert_same(pl[b, :k], o[1]); _assert_same(ep[b, :k], o[2])
        _assert_same(tm[b, :k], o[3]); _assert_same(rs[b, :k], o[4]); _assert_same(vs[b, :k], o[5])


def test_edge_cases(engines, oracle, systems):
    spec = systems["doc2"]; eng = engines["doc2"]; osys = oracle.OracleSystem(spec)
    rng = np.random.default_rng(9)
    # ragged batch sizes around the tile (128) and warp (32) boundaries, and B=1
    for B in (1, 31, 32, 33, 127, 128, 129, 257):
        X = rng.random((B, spec["ndim"]))
        chi2, logl, st = eng.loglike(X, mode=0)
        o = osys.loglike(X, mode=0)
        _assert_same(chi2, o[0]); _assert_same(logl, o[1]); _assert_same(st, o[2])
    # empty batch
    chi2, logl, st = eng.loglike(np.zeros((0, spec["ndim"])), mode=0)
    assert chi2.size == 0
    # out-of-hypercube rows: +inf / -inf like utils.py:245-246 and :368-369, NaN is outside
    X = rng.random((64, spec["ndim"]))
    X[3, 0] = -1e-12; X[7, -1] = 1 + 1e-12; X[11, 2] = np.nan; X[12] = 0.0; X[13] = 1.0
    chi2, logl, st = eng.loglike(X, mode=0)
    o = osys.loglike(X, mode=0)
    for r in (3, 7, 11):
        assert chi2[r] == math.inf and logl[r] == -math.inf and st[r] & 1
    _assert_same(chi2, o[0]); _assert_same(st, o[2])
    # wrong row length is refused on the host side
    with pytest.raises(ValueError):
        eng.loglike(np.zeros((4, spec["ndim"] + 1)), mode=0)


def test_physical_modes(engines, oracle, systems, known):
    spec = systems["sys3"]; eng = engines["sys3"]; osys = oracle.OracleSystem(spec)
    phys = np.array(known["G1"]["phys_free"])
    rng = np.random.default_rng(2)
    X = phys[None, :] * (1 + 1e-6 * rng.standard_normal((40, phys.size)))
    chi2, logl, st = eng.loglike(X, mode=1)
    o = osys.loglike(X, mode=1)
    _assert_same(chi2, o[0]); _assert_same(logl, o[1]); _assert_same(st, o[2])
    assert (st == 0).any()


def test_device_pointer_path_matches_host_path(engines, systems):
    import torch
    spec = systems["sys3"]; eng = engines["sys3"]
    X = _cubes(spec, 777, 4)
    chi2, logl, st = eng.loglike(X, mode=0)
    Xd = torch.from_numpy(X).cuda()
    c2, l2, s2 = eng.loglike(Xd, mode=0)
    torch.cuda.synchronize()
    _assert_same(chi2, c2.cpu().numpy()); _assert_same(logl, l2.cpu().numpy()); _assert_same(st, s2.cpu().numpy())
    assert eng.last_kernel_ms() > 0 and eng.launch_count() >= 2


def test_full_size_properties(engines, systems):
    # BASELINE config 2 at full size (10^4 members): size-independent properties
    spec = systems["doc2"]; eng = engines["doc2"]
    rng = np.random.default_rng(20240607)
    X = rng.random((10000, spec["ndim"]))
    chi2, logl, st = eng.loglike(X, mode=0)
    assert np.all(np.isfinite(chi2)) and np.all(chi2 > 0)
    assert np.array_equal(logl, -0.5 * chi2 - spec["second_term_sum"])
    # permutation invariance: every system is independent of its position in the batch
    perm = rng.permutation(10000)
    chi2p, _, _ = eng.loglike(X[perm], mode=0)
    assert np.array_equal(chi2p, chi2[perm])
    # idempotence: a second call gives identical bits
    chi2b, _, _ = eng.loglike(X, mode=0)
    assert np.array_equal(chi2b, chi2)


def test_c4_full_size_properties(oracle):
    """BASELINE config 4 at full size (3 planets, 1500 d, dt 0.1 -> 15001 steps, 10^5 walkers):
    size-independent properties, plus a sample against the oracle."""
    import bench
    import nauyaca_b200 as nb
    spec, X, _ = bench.make_workload("c4", lambda s: nb.BatchEngine(s))
    eng = nb.BatchEngine(spec)
    assert eng.nsteps == 15001 and X.shape == (100_000, 18)
    chi2, logl, st = eng.loglike(X, mode=0)
    assert np.all(st == 0) and np.all(np.isfinite(chi2)) and np.all(chi2 > 0)
    assert np.array_equal(logl, -0.5 * chi2 - spec["second_term_sum"])
    # the centre of the walker cloud (the true solution projected onto the free parameters; the
    # fixed inclinations/node of the system differ slightly from the truth) fits far better
    # than the cloud around it
    c = bench.cube_of(spec, bench.TRUES_3PL)
    c0, _, _ = eng.loglike(c[None, :], mode=0)
    assert c0[0] < 1e-3 * np.median(chi2)
    # position independence across tiles, warps and time segments
    rng = np.random.default_rng(1)
    idx = rng.permutation(100_000)[:5000]
    chi2s, _, _ = eng.loglike(X[idx], mode=0)
    assert np.array_equal(chi2s, chi2[idx])
    # a sample of 512 walkers against the oracle, bit for bit (host threads: ~10 ms per system)
    osys = oracle.OracleSystem(spec)
    pick = idx[:512]
    oc, ol, ost = osys.loglike_threads(X[pick], mode=0)
    assert np.array_equal(oc, chi2[pick]) and np.array_equal(ol, logl[pick]) and np.all(ost == 0)
    # ... and the whole event table (planet, epoch, time, rsky, vsky) of 32 of them at 15001 steps
    flat = np.stack([osys.cube_to_physical(x) for x in X[pick[:32]]])
    n, pl, ep, tm, rs, vs, st2 = eng.ephemeris(flat, mode=2)
    for b in range(32):
        o = oracle.ephemeris(flat[b], spec["mstar"], spec["t0"], spec["dt"], spec["ftime"])
        k = int(n[b])
        assert o[0] == st2[b] == 0 and k == len(o[3]) and k > 300
        assert np.array_equal(pl[b, :k], o[1]) and np.array_equal(ep[b, :k], o[2])
        assert np.array_equal(tm[b, :k], o[3]) and np.array_equal(rs[b, :k], o[4]) and np.array_equal(vs[b, :k], o[5])
    eng.close()


def test_c5_full_sweep_streams_in_chunks(oracle, systems):
    """BASELINE config 5 at full size: the 10^7-point uniform sweep of the 3-planet system
    (500 d, dt 0.34874) through the host-pointer path, which streams it in chunks of 2^20.
    Size-independent properties (every chunk boundary included) and a sample against the
    oracle, bit for bit."""
    import nauyaca_b200 as nb
    spec = systems["sys3"]
    eng = nb.BatchEngine(spec); osys = oracle.OracleSystem(spec)
    B = 10_000_000
    rng = np.random.default_rng(20240607)
    X = rng.random((B, spec["ndim"]))
    chi2, logl, st = eng.loglike(X, mode=0)
    assert chi2.shape == (B,) and np.all(chi2 > 0) and not np.isnan(chi2).any()
    assert np.all((st & 1) == 0)                                        # U(0,1) walkers are inside the hypercube
    assert np.array_equal(logl, -0.5 * chi2 - spec["second_term_sum"])
    # every system is independent of its position in the batch, chunk boundaries included
    edges = np.concatenate([np.arange(c - 40, c + 40) for c in range(1 << 20, B, 1 << 20)])
    idx = np.concatenate([edges, rng.integers(0, B, 4000), [0, B - 1]])
    chi2s, _, sts = eng.loglike(X[idx], mode=0)
    assert np.array_equal(chi2s, chi2[idx]) and np.array_equal(sts, st[idx])
    pick = idx[::200]
    oc, ol, ost = osys.loglike(X[pick], mode=0)
    assert np.array_equal(oc, chi2[pick]) and np.array_equal(ost, st[pick])
    eng.close()


@pytest.mark.parametrize("dt,key", [(0.34874, "chi2_regression"), (0.1, "chi2_regression_dt0.1")])
def test_g5_long_baseline_on_gpu(oracle, systems, dt, key):
    """SURVEY section 4, G5 on the GPU: the 1359-d baseline (3897 and 13591 steps), 285 events, none masked,
    chi2 78.56 / 44.08 / 30.38, event table and chi2 equal to the oracle's bit for bit."""
    import nauyaca_b200 as nb
    from test_oracle_golden import g5_spec
    g5, spec = g5_spec(systems)
    spec["dt"] = dt
    eng = nb.BatchEngine(spec); osys = oracle.OracleSystem(spec)
    flat = np.array([g5["flat_params"]])
    chi2, logl, st, per = eng.loglike(flat, mode=2, per_planet=True)
    assert st[0] == 0
    for i in range(3):
        assert abs(per[0, i] - g5["chi2_survey_2dec"][i]) <= 0.0051
        assert abs(per[0, i] - g5[key][i]) <= 1e-6 * g5[key][i]
    oc, ol, ost = osys.loglike(flat, mode=2)
    assert chi2[0] == oc[0] and logl[0] == ol[0]
    n, pl, ep, tm, rs, vs, st2 = eng.ephemeris(flat, mode=2)
    o = oracle.ephemeris(flat[0], spec["mstar"], spec["t0"], dt, spec["ftime"])
    k = int(n[0])
    assert k == 285 and [int((pl[0, :k] == i).sum()) for i in range(3)] == g5["n_events"]
    assert np.array_equal(tm[0, :k], o[3]) and np.array_equal(ep[0, :k], o[2]) and np.array_equal(rs[0, :k], o[4])
    assert np.all(rs[0, :k] <= spec["rstarAU"])
    eng.close()


def test_two_body_closed_form_on_gpu():
    """The known answers of tests/test_oracle_properties.py through the C ABI (ttvb_ephemeris, N = 1 and
    N = 2 kernels): two-body transit times against the closed form, no golden and no oracle involved."""
    import nauyaca_b200 as nb
    from test_oracle_properties import CASES, _analytic_transits
    mstar, t0, dt, tend = 0.8, 2.5, 0.05, 120.0
    for mass, P, e, w, M0, Om in CASES:
        SP = nb.utils.run_TTVFast([mass, P, e, 90.0, w, M0, Om], mstar, init_time=t0, final_time=tend, dt=dt)
        tm = SP[2]
        want = _analytic_transits(P, e, w, M0, t0, tend)
        want = want[(want > t0 + dt) & (want < tend - dt)]
        got = tm[(tm > t0 + dt) & (tm < tend - dt)]
        assert len(got) == len(want) and np.max(np.abs(got - want)) < 2e-10
    # massless pair: two independent orbits
    flat = [1e-7, 5.3, 0.1, 90.0, 40.0, 100.0, 0.0, 1e-7, 8.9, 0.2, 90.0, 250.0, 30.0, 0.0]
    SP = nb.utils.run_TTVFast(flat, 1.1, init_time=0.0, final_time=150.0, dt=0.04)
    for j, (P, e, w, M0) in enumerate([(5.3, 0.1, 40.0, 100.0), (8.9, 0.2, 250.0, 30.0)]):
        t = SP[2][SP[0] == j]
        want = _analytic_transits(P, e, w, M0, 0.0, 150.0)
        want = want[(want > 0.04) & (want < 150.0 - 0.04)]
        t = t[(t > 0.04) & (t < 150.0 - 0.04)]
        assert len(t) == len(want) and np.max(np.abs(t - want)) < 5e-9
