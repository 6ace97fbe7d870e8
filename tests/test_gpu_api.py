"""GPU: the host-side mirror of the reference interface (nauyaca_b200.utils, GpuPool, DE seam)
through the C ABI, against the goldens and the oracle."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_calculate_ephemeris_golden(systems, known, spec_ps):
    from nauyaca_b200 import utils as U
    PS = spec_ps(systems["sys3"])
    for g in ("G1", "G2"):
        eph = U.calculate_ephemeris(PS, known[g]["flat_params"])
        for pid, ge in known[g]["ephemeris"].items():
            assert list(eph[pid].keys()) == ge["epochs"]
            if ge["times"]:
                assert np.max(np.abs(np.array(list(eph[pid].values())) - np.array(ge["times"]))) <= 1e-7
    SP = U.run_TTVFast(known["G1"]["flat_params"], PS.mstar, init_time=PS.t0, final_time=PS.ftime, dt=PS.dt)
    assert SP.shape == (5, 105)


def test_objectives_match_reference_semantics(systems, known, spec_ps, oracle):
    from nauyaca_b200 import utils as U
    spec = systems["sys3"]; PS = spec_ps(spec); osys = oracle.OracleSystem(spec)
    k = known["G1"]
    # physical + insert_constants, individual chi2 (miscellany.ipynb cells 20-23, sigma convention /2)
    ind = U.calculate_chi2_physical(k["phys_free"], PS, insert_constants=True, individual=True)
    for pid, v in k["chi2_individual_old_sigma"].items():
        assert abs(ind[pid] - v / 2) <= 2e-9 * v / 2
    chi2 = U.calculate_chi2_physical(k["phys_free"], PS, insert_constants=True)
    assert abs(chi2 - k["chi2_current"]) <= 1e-9 * k["chi2_current"]
    ll = U.log_likelihood_func(k["phys_free"], PS, cube=False, insert_constants=True)
    assert ll == -0.5 * chi2 - sum(PS.second_term_logL.values())
    chi2b, eph = U.calculate_chi2_physical(k["flat_params"], PS, get_ephemeris=True)
    assert chi2b == chi2 and len(eph["Planet-b"]) == 48
    # G2: 36 missing epochs -> 3.6e21, and the notebook's -inf for the full vector with
    # cube=False (bounds check misaligned, miscellany.ipynb cells 49-51)
    assert U.calculate_chi2_physical(known["G2"]["flat_params"], PS) == 3.6e21
    assert U.log_likelihood_func(known["G2"]["flat_params"], PS, cube=False) == -np.inf
    with pytest.raises(SystemExit):
        U.calculate_chi2_physical(k["phys_free"], PS)           # utils.py:309-312
    # cube objective and bounds
    x = np.random.default_rng(0).random(spec["ndim"])
    assert U.calculate_chi2(x, PS) == osys.loglike(x[None, :], mode=0)[0][0]
    assert U.log_likelihood_func(x, PS) == osys.loglike(x[None, :], mode=0)[1][0]
    x[3] = 1.0000001
    assert U.calculate_chi2(x, PS) == np.inf and U.log_likelihood_func(x, PS) == -np.inf
    assert list(U.cube_to_physical(PS, np.full(spec["ndim"], 0.25))) == list(osys.cube_to_physical(np.full(spec["ndim"], 0.25)))


def test_shrunken_hypercube_is_followed(systems, spec_ps):
    # optimizers.py:90 mutates PSystem.hypercube in place
    from nauyaca_b200 import utils as U
    PS = spec_ps(systems["doc2"])
    x = np.full(PS.ndim, 0.5)
    assert np.isfinite(U.calculate_chi2(x, PS))
    PS.hypercube = [[0.6, 0.9]] * PS.ndim
    assert U.calculate_chi2(x, PS) == np.inf
    PS.hypercube = [[0.0, 1.0]] * PS.ndim
    assert np.isfinite(U.calculate_chi2(x, PS))


def test_gpupool_and_de_seam_on_gpu(systems, spec_ps, oracle):
    import nauyaca_b200 as nb
    from nauyaca_b200 import utils as U
    from test_host_logic import _Evaluator
    spec = systems["doc2"]; PS = spec_ps(spec); osys = oracle.OracleSystem(spec)
    rng = np.random.default_rng(8)
    P = rng.random((16 * 11, spec["ndim"]))                    # a ptemcee half-step block
    ev = _Evaluator(U.log_likelihood_func, lambda x: 0.0, [PS, True, False])
    res = nb.GpuPool().map(ev, list(P))
    ref = osys.loglike(P, mode=0)[1]
    assert [r[0] for r in res] == list(ref) and all(r[1] == 0.0 for r in res)
    # scipy differential evolution, one vectorised generation (optimizers.py:78-83 settings)
    from scipy.optimize import differential_evolution
    seen = []
    f = nb.vectorized_chi2(PS)

    def obj(x):
        out = f(x)
        seen.append(np.atleast_1d(out).size)
        return out
    r = differential_evolution(obj, PS.hypercube, popsize=30, maxiter=1, mutation=(1.2, 1.7), recombination=0.6,
                               polish=False, vectorized=True, updating="deferred", seed=1, tol=0.0)
    assert max(seen) == 30 * spec["ndim"]                       # the whole population in one batch
    assert np.isfinite(r.fun) and r.fun == osys.loglike(r.x[None, :], mode=0)[0][0]


def test_pt_sampler_on_gpu(systems, known, spec_ps, oracle):
    """mcmc.py:195-221 on the GPU path: pool.map route == block route, walkers climb."""
    import nauyaca_b200 as nb
    spec = systems["doc2"]; PS = spec_ps(spec); osys = oracle.OracleSystem(spec)
    import bench
    c = bench.cube_of(spec, known["G4"]["flat_params"])
    rng = np.random.mtrand.RandomState(3)
    nt, nw = 4, 2 * spec["ndim"] + 2
    p0 = np.clip(c[None, None, :] + 0.02 * rng.normal(size=(nt, nw, spec["ndim"])), 0, 1)
    s1 = nb.mcmc.make_sampler(PS, p0, tmax=50.0, random=np.random.mtrand.RandomState(5))
    s2 = nb.mcmc.make_sampler(PS, p0, tmax=50.0, random=np.random.mtrand.RandomState(5), use_pool=True)
    r1 = s1.run_mcmc(p0, iterations=30, adapt=True)
    r2 = s2.run_mcmc(p0, iterations=30, adapt=True)
    assert np.array_equal(r1[0], r2[0]) and np.array_equal(r1[2], r2[2])
    # the stored log-likelihoods are the oracle's, bit for bit
    ref = osys.loglike(r1[0].reshape(-1, spec["ndim"]), mode=0)[1].reshape(nt, nw)
    assert np.array_equal(r1[2], ref)
    l0 = osys.loglike(p0[0], mode=0)[1]
    assert np.median(r1[2][0]) > np.median(l0)


def test_batched_optimizer_chains_on_gpu(systems, spec_ps, oracle):
    """optimizers.py:46-139 seam: DE -> Powell -> Nelder-Mead chains in rendezvous batches."""
    import nauyaca_b200 as nb
    spec = systems["doc2"]; PS = spec_ps(spec); osys = oracle.OracleSystem(spec)
    out = nb.optimizers.run_optimizers(PS, nsols=8, seed=7, powell_opts={"maxfev": 300}, nm_opts={"maxfev": 500},
                                       greedy=False)
    assert out["chi2"].shape == (8,) and np.all(np.diff(out["chi2"]) >= 0)
    # every reported chi2 is the oracle's chi2 of the reported cube vector, bit for bit
    ref = osys.loglike(out["cube"], mode=0)[0]
    assert np.array_equal(out["chi2"], ref)
    st = out["info"]["stages"]
    assert all(s[2] <= s[1] <= s[0] for s in st)                        # each stage improves
    assert max(out["info"]["batches"]) == 8 * 30 * spec["ndim"]         # a DE generation of all chains in one launch
    assert out["physical"].shape == (8, spec["ndim"])
    # the default (greedy) batching policy changes how the evaluations are grouped, not what they return
    g = nb.optimizers.run_optimizers(PS, nsols=8, seed=7, powell_opts={"maxfev": 300}, nm_opts={"maxfev": 500})
    assert np.array_equal(g["chi2"], out["chi2"]) and np.array_equal(g["cube"], out["cube"])
    assert sum(g["info"]["batches"]) == sum(out["info"]["batches"])


def test_device_sampler(systems, known, spec_ps, oracle):
    """ttvb_pt_*: the PT iteration on the device. Bookkeeping invariants (every stored logl is
    the oracle's logl of the stored walker, bit for bit; the ladder keeps its ends and its
    order; counters add up) and statistical agreement with the host sampler."""
    import nauyaca_b200 as nb
    import bench
    spec = systems["doc2"]; PS = spec_ps(spec); osys = oracle.OracleSystem(spec)
    c = bench.cube_of(spec, known["G4"]["flat_params"])
    rng = np.random.mtrand.RandomState(11)
    nt, nw, dim = 6, 60, spec["ndim"]
    p0 = np.clip(c[None, None, :] + 0.01 * rng.normal(size=(nt, nw, dim)), 0, 1)
    dev = nb.mcmc.make_device_sampler(PS, p0, tmax=50.0, seed=1234)
    n = 0
    for p, logpost, logl in dev.sample(p0, iterations=40, thin=4, adapt=True):
        n += 1
    assert n == 40 and dev.chain.shape == (nt, nw, 10, dim)
    ref = osys.loglike(p.reshape(-1, dim), mode=0)[1].reshape(nt, nw)
    assert np.array_equal(logl, ref)                                   # p and logl moved together through accepts and swaps
    assert np.allclose(logpost, logl * dev.betas[:, None])
    b = dev.betas
    assert b[0] == 1.0 and abs(b[-1] - 1 / 50.0) < 1e-15 and np.all(np.diff(b) < 0)
    assert np.all(dev.nprop == 40) and np.all(dev.nprop_accepted <= 40)
    assert dev.nswap[0] == 40 * nw and dev.nswap[1] == 80 * nw and np.all(dev.nswap_accepted <= dev.nswap)
    assert 0.05 < dev.acceptance_fraction[0].mean() < 0.95
    assert dev.get_autocorr_time().shape == (nt, dim)
    # same seed -> same chain (counter-based RNG), different seed -> different chain
    dev2 = nb.mcmc.make_device_sampler(PS, p0, tmax=50.0, seed=1234)
    r2 = dev2.run_mcmc(p0, iterations=40, adapt=True)
    assert np.array_equal(r2[0], p)
    # advance(n) == n single iterations
    dev3 = nb.mcmc.make_device_sampler(PS, p0, tmax=50.0, seed=1234)
    next(dev3.sample(p0, iterations=1, adapt=True, storechain=False))
    r3 = dev3.advance(39, adapt=True)
    assert np.array_equal(r3[0], p)
    # statistics against the host sampler after the same number of iterations
    host = nb.mcmc.make_sampler(PS, p0, tmax=50.0, random=np.random.mtrand.RandomState(5))
    rh = host.run_mcmc(p0, iterations=150, adapt=True)
    rd = nb.mcmc.make_device_sampler(PS, p0, tmax=50.0, seed=77)
    next(rd.sample(p0, iterations=1, adapt=True, storechain=False))
    pd_, lpd, lld = rd.advance(149, adapt=True)
    assert abs(np.median(lld[0]) - np.median(rh[2][0])) < 6.0        # cold chains reach the same logl level
    assert abs(rd.acceptance_fraction[0].mean() - host.acceptance_fraction[0].mean()) < 0.12
    for s in (dev, dev2, dev3, rd):
        s.close()


def test_device_sampler_sharded(systems, known, spec_ps, tmp_path):
    """The device sampler with the proposals of every half-step sharded over two ranks and the
    proposal log-likelihoods all-gathered (ttvb_pt_half_begin / _end): both ranks end with the
    ensemble of the unsharded sampler, bit for bit."""
    import subprocess
    import sys
    import nauyaca_b200 as nb
    import bench
    spec = systems["doc2"]; PS = spec_ps(spec)
    c = bench.cube_of(spec, known["G4"]["flat_params"])
    rng = np.random.mtrand.RandomState(11)
    nt, nw, dim = 5, 46, spec["ndim"]              # 115 proposals per half-step: uneven shards (58 + 57)
    p0 = np.clip(c[None, None, :] + 0.01 * rng.normal(size=(nt, nw, dim)), 0, 1)
    one = nb.mcmc.make_device_sampler(PS, p0, tmax=50.0, seed=1234)
    next(one.sample(p0, iterations=1, adapt=True, storechain=False))
    p, logpost, logl = one.advance(24, adapt=True)
    np.save(tmp_path / "p0.npy", p0)
    port = str(29600 + os.getpid() % 300)
    worker = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_shard_sampler_worker.py")
    procs = [subprocess.Popen([sys.executable, worker, str(r), "2", port, str(tmp_path / f"r{r}.npz"),
                               str(tmp_path / "p0.npy")]) for r in range(2)]
    for pr in procs:
        assert pr.wait(timeout=300) == 0
    for r in range(2):
        z = np.load(tmp_path / f"r{r}.npz")
        assert np.array_equal(z["p"], p) and np.array_equal(z["logl"], logl)
        assert np.array_equal(z["betas"], one.betas) and np.array_equal(z["nacc"], one.nprop_accepted)
    # without a seed: rank 0 draws one and broadcasts it, so both ranks still hold ONE ensemble
    procs = [subprocess.Popen([sys.executable, worker, str(r), "2", str(int(port) + 1), str(tmp_path / f"u{r}.npz"),
                               str(tmp_path / "p0.npy"), "gloo", "none"]) for r in range(2)]
    for pr in procs:
        assert pr.wait(timeout=300) == 0
    u0, u1 = np.load(tmp_path / "u0.npz"), np.load(tmp_path / "u1.npz")
    assert np.array_equal(u0["p"], u1["p"]) and np.array_equal(u0["logl"], u1["logl"])
    assert np.isfinite(u0["logl"]).all() and not np.array_equal(u0["p"], p)
    # NCCL path (one GPU per rank, in-place all-gather of the proposal log-likelihoods on the
    # device buffer): needs two GPUs
    if nb._lib.lib().ttvb_device_count() >= 2:
        procs = [subprocess.Popen([sys.executable, worker, str(r), "2", str(int(port) + 2), str(tmp_path / f"n{r}.npz"),
                                   str(tmp_path / "p0.npy"), "nccl"]) for r in range(2)]
        for pr in procs:
            assert pr.wait(timeout=300) == 0
        for r in range(2):
            z = np.load(tmp_path / f"n{r}.npz")
            assert np.array_equal(z["p"], p) and np.array_equal(z["logl"], logl)
            assert np.array_equal(z["betas"], one.betas) and np.array_equal(z["nacc"], one.nprop_accepted)
    one.close()


def test_notebook_flow_end_to_end(capsys):
    """examples/fit_two_planets.py: optimizer chains -> init_walkers('Ladder') -> device PT MCMC on
    the Documentation system; the chain must not lose the optimizers' best solution."""
    import importlib.util
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples", "fit_two_planets.py")
    spec = importlib.util.spec_from_file_location("fit_two_planets", path)
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    r = mod.fit(nsols=6, ntemps=4, nwalkers=40, iterations=40, seed=5, quiet=True)
    capsys.readouterr()
    assert r["p0"].shape == (4, 40, 11) and np.isfinite(r["logl"]).all()
    assert np.isfinite(r["best_chi2"]) and r["best_chi2"] < 1e19            # every observed transit reproduced
    assert r["best_chi2"] <= 1.5 * r["opt_chi2"][0] + 50.0


def test_run_mcmc_device_sampler(systems, known, spec_ps, oracle, tmp_path):
    """mcmc.run_mcmc (the reference's MCMC.run bookkeeping) over the device-resident sampler: the
    record holds the ensemble at every intra_steps-th iteration of the same-seed sampler."""
    import nauyaca_b200 as nb
    import bench
    spec = systems["doc2"]; PS = spec_ps(spec); osys = oracle.OracleSystem(spec)
    c = bench.cube_of(spec, known["G4"]["flat_params"])
    rng = np.random.mtrand.RandomState(2)
    nt, nw, dim = 4, 30, spec["ndim"]
    p0 = np.clip(c[None, None, :] + 0.01 * rng.normal(size=(nt, nw, dim)), 0, 1)
    s, rec = nb.mcmc.run_mcmc(PS, p0, itmax=14, intra_steps=4, tmax=30.0, seed=99, path=str(tmp_path) + "/",
                              file_name="dev")
    assert rec["CHAINS"].shape == (nt, nw, 4, dim) and rec["INDEX"][0] == 2 and rec["ITER_LAST"][0] == 12
    ref = nb.mcmc.make_device_sampler(PS, p0, tmax=30.0, seed=99)
    next(ref.sample(p0, iterations=1, adapt=True, storechain=False))
    state = ref.advance(3, adapt=True)
    for k in range(3):
        p, lp, ll = state
        assert np.array_equal(rec["CHAINS"][:, :, k, :], p)
        assert rec["MAP"][k] == lp.max() and rec["MEANLOGPOST"][k] == np.mean(lp[0])
        i, j = np.unravel_index(np.flatnonzero(lp.ravel() == lp.max())[-1], lp.shape)
        assert np.array_equal(rec["BESTSOLS"][k], p[i][j])
        assert np.array_equal(rec["BETAS"][k], ref.betas)
        assert osys.loglike(rec["BESTSOLS"][k][None, :], mode=0)[1][0] * ref.betas[i] == rec["MAP"][k]
        state = ref.advance(4, adapt=True)
    assert np.all(rec["CHAINS"][:, :, 3, :] == 0)                      # the slot the reference leaves empty
    assert os.path.exists(str(tmp_path) + "/dev.npz")
    # the .best table written at the end of the run (mcmc.py:304-305), physical vectors from the GPU
    ranked = nb.mcmc.extract_best_solutions(PS, str(tmp_path) + "/dev.npz", write_file=False)
    lines = open(str(tmp_path) + "/dev.best").read().splitlines()
    assert len(lines) == 3 + len(ranked) and float(lines[3].split()[0]) == ranked[0][0] == rec["MAP"][:3].max()
    k0 = int(np.flatnonzero(rec["MAP"][:3] == ranked[0][0])[-1])
    assert list(ranked[0][1]) == list(osys.cube_to_physical(rec["BESTSOLS"][k0]))
    s.close(); ref.close()
