#!/usr/bin/env python
"""The flow of the reference's Documentation notebooks (2- Setup, 3- Running Optimizers,
4- Running MCMC) on the GPU engine: optimizer chains -> initial walkers -> parallel-tempering
MCMC -> best solution, for the Documentation two-planet system.

    python examples/fit_two_planets.py [--nsols 16] [--ntemps 10] [--nwalkers 200] [--iterations 300]

With nauyaca installed, build the system with nauyaca's own SetPlanet / PlanetarySystem and pass
that object wherever ``PS`` is used below; here the frozen attributes of the notebook's system
(tests/golden/systems.json, written by tests/golden/make_golden.py) stand in for it.
Needs a CUDA device (there is no CPU fallback).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nauyaca_b200 as nb  # noqa: E402


def fit(nsols=16, ntemps=10, nwalkers=200, iterations=300, seed=3, quiet=False):
    say = (lambda *a: None) if quiet else print
    with open(os.path.join(ROOT, "tests", "golden", "systems.json")) as fh:
        PS = nb.SpecSystem(json.load(fh)["doc2"])
    # ---- 3- Running Optimizers: Optimizers(PS, nsols=...).run()  (optimizers.py:142-203)
    t0 = time.perf_counter()
    opt = nb.optimizers.run_optimizers(PS, nsols=nsols, seed=seed)
    say(f"optimizers: {nsols} DE->Powell->Nelder-Mead chains in {time.perf_counter() - t0:.1f} s, "
        f"best chi2 {opt['chi2'][0]:.3f}, largest batch {max(opt['info']['batches'])}")
    # ---- 4- Running MCMC: p0 = utils.init_walkers(PS, distribution='ladder', opt_data=...)  (utils.py:463)
    np.random.seed(seed)
    p0 = nb.utils.init_walkers(PS, distribution="Ladder", opt_data={"chi2": opt["chi2"], "cube": opt["cube"]},
                               ntemps=ntemps, nwalkers=nwalkers, fbest=0.5)
    # ---- MCMC(PS, p0=p0, tmax=..., itmax=..., intra_steps=...).run()  (mcmc.py:141-312), sampler moves
    #      on the device; rec is the record the reference keeps in its HDF5 file (same names)
    intra = max(1, iterations // 10)
    t0 = time.perf_counter()
    sampler, rec = nb.mcmc.run_mcmc(PS, p0, itmax=iterations, intra_steps=intra, tmax=100.0, seed=seed, save=False)
    dt = time.perf_counter() - t0
    p, logpost, logl = sampler._fetch()
    k = int(np.argmax(rec["MAP"][:rec["INDEX"][0] + 1]))
    best = rec["BESTSOLS"][k]
    say(f"mcmc: {int(rec['ITER_LAST'][0])} iterations of {ntemps} x {nwalkers} walkers in {dt:.2f} s "
        f"({int(rec['ITER_LAST'][0]) * ntemps * nwalkers / dt:.0f} likelihood evaluations per second), "
        f"cold-chain acceptance {sampler.acceptance_fraction[0].mean():.2f}, swaps {sampler.tswap_acceptance_fraction[0]:.2f}, "
        f"mean log-posterior of the cold chain {rec['MEANLOGPOST'][0]:.1f} -> {rec['MEANLOGPOST'][rec['INDEX'][0]]:.1f}")
    chi2 = nb.utils.calculate_chi2(best, PS)
    phys = nb.utils.cube_to_physical(PS, best)
    say(f"maximum a posteriori over the run: log-posterior {rec['MAP'][k]:.3f}, chi2 {chi2:.3f}")
    say("physical parameters (mass[Mearth] period[d] ecc inc[deg] argument[deg] mean_anomaly[deg] ascending_node[deg]):")
    for pl in range(PS.npla):
        say("   ", np.array2string(np.asarray(phys[7 * pl:7 * pl + 7]), precision=5))
    sampler.close()
    return {"opt_chi2": opt["chi2"], "best_chi2": float(chi2), "best_logl": float(rec["MAP"][k]), "p0": p0, "logl": logl,
            "record": rec}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--nsols", type=int, default=16)
    ap.add_argument("--ntemps", type=int, default=10)
    ap.add_argument("--nwalkers", type=int, default=200)
    ap.add_argument("--iterations", type=int, default=300)
    a = ap.parse_args()
    fit(a.nsols, a.ntemps, a.nwalkers, a.iterations)
