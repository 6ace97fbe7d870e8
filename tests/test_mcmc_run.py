"""The bookkeeping of the reference's MCMC.run loop (mcmc.py:218-312, 315-402) in
nauyaca_b200.mcmc.run_mcmc: CPU test with the host sampler over the oracle; the device sampler
is covered in tests/test_gpu_api.py."""
import json
import os

import numpy as np
import pytest

import nauyaca_b200 as nb
from conftest import GOLDEN


def _setup(oracle):
    with open(os.path.join(GOLDEN, "systems.json")) as fh:
        spec = json.load(fh)["doc2"]
    PS = nb.SpecSystem(spec)
    osys = oracle.OracleSystem(spec)
    rng = np.random.mtrand.RandomState(3)
    p0 = np.clip(0.5 + 0.05 * rng.normal(size=(3, 24, spec["ndim"])), 0, 1)
    mk = lambda: nb.mcmc.make_sampler(PS, p0, tmax=20.0, random=np.random.mtrand.RandomState(4),
                                      evaluate=lambda X: osys.loglike(X, mode=0)[1])
    return spec, PS, osys, p0, mk


def test_record_follows_the_reference_loop(oracle, tmp_path, capsys):
    spec, PS, osys, p0, mk = _setup(oracle)
    itmax, intra = 13, 3
    to_phys = lambda X: np.array([osys.cube_to_physical(x) for x in X])
    s, rec = nb.mcmc.run_mcmc(PS, p0, itmax=itmax, intra_steps=intra, tmax=20.0, sampler=mk(),
                              path=str(tmp_path) + "/", file_name="run", suffix="_a", to_physical=to_phys)
    nsteps = -(-itmax // intra)
    nlooks = itmax // intra
    assert rec["CHAINS"].shape == (3, 24, nsteps, spec["ndim"]) and rec["BETAS"].shape == (nsteps, 3)
    assert rec["INDEX"][0] == nlooks - 1 and rec["ITER_LAST"][0] == nlooks * intra
    # the same loop by hand: sampler.sample(thin=intra_steps), a look every intra_steps iterations
    ref = mk()
    k = 0
    for it, (p, lp, ll) in enumerate(ref.sample(p0=p0, iterations=itmax, thin=intra, storechain=True, adapt=True)):
        if (it + 1) % intra:
            continue
        assert np.array_equal(rec["CHAINS"][:, :, k, :], p) and np.array_equal(rec["CHAINS"][:, :, k, :], ref.chain[:, :, k, :])
        mx, (i, j) = max((x, (i, j)) for i, row in enumerate(lp) for j, x in enumerate(row))     # mcmc.py:235-237
        assert rec["MAP"][k] == mx and np.array_equal(rec["BESTSOLS"][k], p[i][j])
        assert rec["MEANLOGPOST"][k] == np.mean(lp[0])
        assert np.array_equal(rec["BETAS"][k], ref.betas)
        assert np.array_equal(rec["ACC_FRAC0"][k], ref.tswap_acceptance_fraction)
        assert rec["AUTOCORR"][k] == np.mean(ref.get_autocorr_time()[0])
        k += 1
    assert k == nlooks
    assert np.all(rec["CHAINS"][:, :, nlooks:, :] == 0)                  # the slot the reference leaves empty
    # every stored walker with its stored MAP: logpost = beta * logl (flat prior)
    ll = osys.loglike(rec["BESTSOLS"][:nlooks], mode=0)[1]
    assert np.all(np.isfinite(ll))
    # constants and names
    assert rec["NDIM"][0] == spec["ndim"] and rec["NTEMPS"][0] == 3 and rec["NWALKERS"][0] == 24
    assert rec["ITMAX"][0] == itmax and rec["INTRA_STEPS"][0] == intra and rec["NPLA"][0] == 2
    assert np.array_equal(rec["p0"], p0) and np.array_equal(rec["BOUNDS"], np.array(spec["bounds"]))
    # the file
    z = np.load(str(tmp_path) + "/run_a.npz")
    for key in ("CHAINS", "AUTOCORR", "BETAS", "ACC_FRAC0", "INDEX", "ITER_LAST", "BESTSOLS", "MAP", "MEANLOGPOST",
                "p0", "BOUNDS", "NDIM", "NTEMPS", "NWALKERS", "ITMAX", "INTRA_STEPS", "REF_EPOCH", "NPLA", "MSTAR", "RSTAR"):
        assert np.array_equal(z[key], rec[key], equal_nan=True), key
    assert str(z["NAME"]) == rec["NAME"]
    # the .best table (utils.py:874-934): best first, physical vectors of the saved solutions
    ranked = nb.mcmc.extract_best_solutions(PS, str(tmp_path) + "/run_a.npz", write_file=False, to_physical=to_phys)
    maps = [m for m, _ in ranked]
    assert maps == sorted(set(rec["MAP"][:nlooks]), reverse=True)
    k0 = int(np.flatnonzero(rec["MAP"][:nlooks] == maps[0])[-1])
    assert np.array_equal(ranked[0][1], osys.cube_to_physical(rec["BESTSOLS"][k0]))
    lines = open(str(tmp_path) + "/run_a.best").read().splitlines()
    assert lines[0].split() == ["#Mstar[Msun]", "Rstar[Rsun]", "Nplanets"] and lines[2].startswith("#MAP")
    assert len(lines) == 3 + len(ranked)
    row = lines[3].split()
    assert float(row[0]) == maps[0] and [float(v) for v in row[1:]] == list(ranked[0][1])


def test_argument_checks(oracle, tmp_path, capsys):
    spec, PS, osys, p0, mk = _setup(oracle)
    with pytest.raises(RuntimeError, match="nwalkers have to be >= 22"):
        nb.mcmc.run_mcmc(PS, p0[:, :20], sampler=mk(), save=False)
    with pytest.raises(RuntimeError, match="differs from that in 'p0'"):
        nb.mcmc.run_mcmc(PS, p0[:, :, :10], sampler=mk(), save=False)
    with pytest.raises(RuntimeError, match="differs from number of temperatures"):
        nb.mcmc.run_mcmc(PS, p0, betas=[1.0, 0.5], sampler=mk(), save=False)
    with pytest.raises(RuntimeError, match="does not exist"):
        nb.mcmc.run_mcmc(PS, p0, sampler=mk(), path=str(tmp_path) + "/nowhere/")
    capsys.readouterr()


def test_restart_continues_from_the_last_saved_ensemble(oracle, tmp_path, capsys):
    spec, PS, osys, p0, mk = _setup(oracle)
    path = str(tmp_path) + "/"
    to_phys = lambda X: np.array([osys.cube_to_physical(x) for x in X])
    s, rec = nb.mcmc.run_mcmc(PS, p0, itmax=9, intra_steps=3, tmax=20.0, sampler=mk(), path=path, file_name="first",
                              to_physical=to_phys)
    args = nb.mcmc.restart_mcmc(PS, "first.npz", in_path=path, out_path=str(tmp_path), itmax=6, intra_steps=2)
    k = rec["INDEX"][0]
    assert np.array_equal(args["p0"], rec["CHAINS"][:, :, k, :]) and np.array_equal(args["betas"], rec["BETAS"][k])
    assert args["file_name"] == "first" and args["suffix"] == "_2" and args["path"] == path
    again = nb.mcmc.restart_mcmc(PS, "first.npz", in_path=path, out_path=path, restart_ladder=True)
    assert np.array_equal(again["betas"], rec["BETAS"][0])
    smp = nb.mcmc.make_sampler(PS, args["p0"], betas=args["betas"], random=np.random.mtrand.RandomState(8),
                               evaluate=lambda X: osys.loglike(X, mode=0)[1])
    s2, rec2 = nb.mcmc.run_mcmc(PS, sampler=smp, to_physical=to_phys, **args)
    assert os.path.exists(path + "first_2.npz") and rec2["ITER_LAST"][0] == 6 and np.array_equal(rec2["p0"], args["p0"])
    with pytest.raises(AssertionError):
        nb.mcmc.restart_mcmc(PS, "first.npz", in_path=path, out_path=path, suffix="")
    with pytest.raises(RuntimeError, match="does not exists"):
        nb.mcmc.restart_mcmc(PS, "first.npz", in_path=path, out_path=path + "nowhere")
    capsys.readouterr()
