"""The sampler seam of the reference's MCMC driver (mcmc.py:97-221) on the batched GPU path.

Only the part of ``MCMC.run`` that feeds the hot path is mirrored: the choice between cube /
physical walkers (mcmc.py:129-137), the ptemcee settings (mcmc.py:182-211: a = 10,
adaptation_lag = 1000, adaptation_time = 100, adapt = True) and the ``sampler.sample`` loop.
``run_mcmc`` adds the bookkeeping of that loop (mcmc.py:218-312, 315-402): the record the reference
keeps in its HDF5 file, under the same dataset names, written as ``.npz`` and -- where h5py is
installed -- as the reference's ``.hdf5`` layout. Printing, restart and the convergence
diagnostics stay with the reference (out of scope)."""
import numpy as np

from . import utils as _u
from .ptsampler import Sampler

__all__ = ["make_sampler", "make_device_sampler", "pt_sample", "run_mcmc", "restart_mcmc", "write_hdf5", "extract_best_solutions"]


def _logprior_zero(x, psystem=None):
    """MCMC.logprior of the reference (mcmc.py:505-509): flat, 0.0."""
    return 0.0


def make_sampler(PSystem, p0, betas=None, tmax=None, logprior=None, evaluate=None, random=None, use_pool=False):
    """ptemcee-compatible sampler wired to the GPU likelihood.

    p0: (ntemps, nwalkers, ndim). If every value lies in [0, 1] the walkers are hypercube
    vectors (cube=True), otherwise physical vectors without the constants
    (cube=False, insert_constants=True), as mcmc.py:129-137 decides.
    evaluate: optional callable X[B, ndim] -> logl[B] replacing the single-GPU engine, e.g. a
    nauyaca_b200.ShardedEvaluator-based function for several GPUs.
    use_pool=True goes through ``pool.map`` (GpuPool) exactly as ptemcee would; the default
    hands whole blocks to the engine."""
    p0 = np.asarray(p0, dtype=float)
    ntemps, nwalkers, ndim = p0.shape
    if ndim != PSystem.ndim:
        raise ValueError("p0 has the wrong number of dimensions")
    p0_norm = bool(((p0 >= 0.0) & (p0 <= 1.0)).all())
    insert_cnst = not p0_norm
    logp = _logprior_zero if logprior is None else logprior
    kw = dict(nwalkers=nwalkers, dim=ndim, logp=logp, logl=_u.log_likelihood_func, ntemps=ntemps, betas=betas,
              adaptation_lag=1000.0, adaptation_time=100.0, a=10, Tmax=tmax,
              loglargs=(PSystem, p0_norm, insert_cnst), logpkwargs={'psystem': PSystem}, random=random)
    if use_pool:
        from .pool import GpuPool
        return Sampler(pool=GpuPool(PSystem), **kw)
    if evaluate is None:
        def evaluate(X):
            return _u.log_likelihood_batch(X, PSystem, cube=p0_norm, insert_constants=insert_cnst)
    # the reference's default prior is identically 0.0 (mcmc.py:505-509): no per-vector calls
    blp = (lambda X: np.zeros(X.shape[0])) if logprior is None else None
    return Sampler(batch_logl=evaluate, batch_logp=blp, **kw)


def make_device_sampler(PSystem, p0, betas=None, tmax=None, seed=None, device=None, shard=False, group=None):
    """The same set-up with every sampler move on the GPU (flat prior only): see
    ptsampler.DeviceSampler. The per-iteration host work disappears; use ``advance(n)`` to run
    n iterations between two looks at the state. shard=True (one process per GPU, torch.distributed
    initialised, same seed and p0 on every rank): the proposals of every half-step are split over
    the ranks and their log-likelihoods all-gathered."""
    from .ptsampler import DeviceSampler
    p0 = np.asarray(p0, dtype=float)
    ntemps, nwalkers, ndim = p0.shape
    if ndim != PSystem.ndim:
        raise ValueError("p0 has the wrong number of dimensions")
    p0_norm = bool(((p0 >= 0.0) & (p0 <= 1.0)).all())
    eng = _u.engine_for(PSystem, device=device)
    return DeviceSampler(eng, nwalkers, ntemps=ntemps, Tmax=tmax, betas=betas, a=10, adaptation_lag=1000.0,
                         adaptation_time=100.0, seed=seed, cube=p0_norm, shard=shard, group=group)


def pt_sample(PSystem, p0, iterations, thin=1, **kw):
    """Generator over sampler.sample(...) with the reference's settings (mcmc.py:218-221);
    yields (sampler, p, logpost, logl) every iteration."""
    sampler = make_sampler(PSystem, p0, **kw)
    for p, logpost, logl in sampler.sample(p0=p0, iterations=iterations, thin=thin, storechain=True, adapt=True,
                                           swap_ratios=False):
        yield sampler, p, logpost, logl


def _attr(PS, name, default):
    v = getattr(PS, name, default)
    return default if v is None else v


def run_mcmc(PSystem, p0, itmax=100, intra_steps=1, tmax=None, betas=None, seed=None, device=None, sampler=None,
             path="./", file_name=None, suffix="", save=True, verbose=False, shard=False, to_physical=None):
    """``MCMC.run`` (mcmc.py:141-312): ``itmax`` parallel-tempering iterations, the state looked
    at every ``intra_steps`` iterations. Returns ``(sampler, record)``; ``record`` holds what the
    reference stores in its HDF5 file, same names, shapes and meaning (mcmc.py:315-402):
    CHAINS (ntemps, nwalkers, nsteps, ndim), AUTOCORR, BETAS, ACC_FRAC0, INDEX, ITER_LAST,
    BESTSOLS, MAP, MEANLOGPOST, p0, BOUNDS, NDIM, NTEMPS, NWALKERS, ITMAX, INTRA_STEPS, REF_EPOCH,
    COL_NAMES, NAME, NPLA, MSTAR, RSTAR.

    sampler: default the device-resident sampler (``make_device_sampler``; its iterations between
    two looks run without any host round trip); a host ``Sampler`` (``make_sampler``, e.g. with a
    Python prior) is driven through ``sample(thin=intra_steps)`` exactly as the reference does.
    save: write ``<path><file_name or system_name><suffix>.npz``, if h5py can be imported the same
    record as ``.hdf5`` in the reference's layout, and the ``.best`` table of
    ``extract_best_solutions``. to_physical: see ``extract_best_solutions``."""
    import os
    p0 = np.ascontiguousarray(p0, dtype=float)
    ntemps, nwalkers, ndim = p0.shape
    if nwalkers < 2 * PSystem.ndim:
        raise RuntimeError(f"Number of walkers must be >= 2*ndim, i.e., nwalkers have to be >= {2 * PSystem.ndim}.")
    if ndim != PSystem.ndim:
        raise RuntimeError(f"Number of dimensions in 'PSystem' ({PSystem.ndim}) differs from that in 'p0' ({ndim}).")
    if betas is not None and len(betas) != ntemps:
        raise RuntimeError(f"Number of 'betas' ({betas}) differs from number of temperatures in 'p0' ({ntemps})")
    if save and not os.path.exists(path):
        raise RuntimeError(f"directory -path- {path} does not exist")
    nsteps = -(-int(itmax) // int(intra_steps))                       # mcmc.py:328
    if sampler is None:
        sampler = make_device_sampler(PSystem, p0, betas=betas, tmax=tmax, seed=seed, device=device, shard=shard)
    rec = {"NAME": _attr(PSystem, "system_name", ""), "COL_NAMES": _attr(PSystem, "params_names", ""),
           "CHAINS": np.zeros((ntemps, nwalkers, nsteps, ndim)), "AUTOCORR": np.zeros(nsteps),
           "BETAS": np.zeros((nsteps, ntemps)), "ACC_FRAC0": np.zeros((nsteps, ntemps)),
           "INDEX": np.zeros(1, np.int64), "ITER_LAST": np.zeros(1, np.int64),
           "BESTSOLS": np.zeros((nsteps, ndim)), "MAP": np.zeros(nsteps), "MEANLOGPOST": np.zeros(nsteps),
           "p0": p0.copy(), "BOUNDS": np.array(_attr(PSystem, "bounds", []), dtype=float),
           "NDIM": np.array([ndim], np.int64), "NTEMPS": np.array([ntemps], np.int64),
           "NWALKERS": np.array([nwalkers], np.int64), "ITMAX": np.array([itmax], np.int64),
           "INTRA_STEPS": np.array([intra_steps], np.int64),
           "REF_EPOCH": np.array([_attr(PSystem, "t0", 0.0)], np.int64),
           "NPLA": np.array([PSystem.npla], np.int64), "MSTAR": np.array([float(PSystem.mstar)]),
           "RSTAR": np.array([float(_attr(PSystem, "rstar", np.nan))])}
    bp = getattr(PSystem, "_bounds_parameterized", None)
    if bp is not None:
        rec["BOUNDSP"] = np.array(bp, dtype=float)
    device_resident = hasattr(sampler, "advance")

    def looks():
        """(iteration, p, logpost, logl) at every intra_steps-th iteration."""
        if device_resident:
            done = 0
            while done + intra_steps <= itmax:
                if done == 0:
                    next(sampler.sample(p0, iterations=1, adapt=True, storechain=False))
                    s = sampler.advance(intra_steps - 1, adapt=True) if intra_steps > 1 else sampler._fetch()
                else:
                    s = sampler.advance(intra_steps, adapt=True)
                done += intra_steps
                yield done - 1, s
        else:
            for it, s in enumerate(sampler.sample(p0=p0, iterations=itmax, thin=intra_steps, storechain=True,
                                                  adapt=True, swap_ratios=False)):
                if (it + 1) % intra_steps:
                    continue
                yield it, s

    index = 0
    for iteration, s in looks():
        p, logpost, logl = s
        # MAP over all temperatures; ties go to the last (temperature, walker), as max() over
        # (value, (i, j)) tuples does in mcmc.py:235-237
        best = np.flatnonzero(logpost.ravel() == logpost.max())[-1]
        bi, bj = np.unravel_index(best, logpost.shape)
        rec["CHAINS"][:, :, index, :] = p
        tau_sampler = sampler
        if device_resident:                                           # its chain lives here, not in the sampler
            tau_sampler = type("_Chain", (), {"ntemps": ntemps, "dim": ndim, "_chain": rec["CHAINS"][:, :, :index + 1, :]})()
        rec["AUTOCORR"][index] = np.mean(Sampler.get_autocorr_time(tau_sampler)[0])
        rec["BETAS"][index, :] = sampler.betas
        rec["ACC_FRAC0"][index, :] = sampler.tswap_acceptance_fraction
        rec["INDEX"][:] = index
        rec["ITER_LAST"][:] = iteration + 1
        rec["BESTSOLS"][index, :] = p[bi][bj]
        rec["MAP"][index] = logpost[bi][bj]
        rec["MEANLOGPOST"][index] = np.mean(logpost[0])
        if verbose:
            print("--------- Iteration: ", iteration + 1)
            print(" Mean tau Temp 0:", round(float(rec["AUTOCORR"][index]), 3))
            print(" Accepted swap fraction in Temp 0: ", round(float(rec["ACC_FRAC0"][index, 0]), 3))
            print(" Mean acceptance fraction Temp 0: ", round(float(np.mean(sampler.acceptance_fraction[0, :])), 3))
            print(" Mean log-likelihood: ", round(float(np.mean(logl[0])), 3))
            print(" Mean log-posterior:  ", round(float(rec["MEANLOGPOST"][index]), 3))
            print(" Current MAP: ", (int(bi), int(bj)), round(float(rec["MAP"][index]), 3))
        index += 1
        if (index + 1) * intra_steps > itmax:
            break
    if save:
        base = f"{path}{file_name if file_name is not None else rec['NAME']}{suffix}"
        np.savez_compressed(base + ".npz", **rec)
        rec["files"] = [base + ".npz"]
        try:
            import h5py  # noqa: F401
        except ImportError:
            pass
        else:
            rec["files"].append(write_hdf5(rec, base + ".hdf5"))
        # mcmc.py:304-305: the best solutions of the run, in physical form, as a text table
        extract_best_solutions(PSystem, rec, write_file=True, best_file=base + ".best", to_physical=to_physical)
        rec["files"].append(base + ".best")
    return sampler, rec


def write_hdf5(rec, filename):
    """The record of ``run_mcmc`` as the reference's HDF5 file (mcmc.py:315-374: same dataset
    names, shapes, dtypes and compression filters), readable by the reference's
    ``extract_best_solutions`` / ``mcmc_summary`` / plots. Needs h5py (not part of this package)."""
    import h5py
    lzf = {"CHAINS"}
    plain = {"INDEX", "ITER_LAST", "BOUNDS", "BOUNDSP", "NDIM", "NTEMPS", "NWALKERS", "ITMAX", "INTRA_STEPS",
             "REF_EPOCH", "NPLA", "MSTAR", "RSTAR", "CORES"}
    with h5py.File(filename, "w") as fh:
        fh["NAME"] = rec["NAME"]
        fh["COL_NAMES"] = rec["COL_NAMES"]
        fh.create_dataset("CORES", (1,), dtype="i8")[:] = 1
        for key, val in rec.items():
            if key in ("NAME", "COL_NAMES", "files"):
                continue
            val = np.asarray(val)
            dt = "i8" if val.dtype.kind == "i" else "f8"
            if key in lzf:
                fh.create_dataset(key, val.shape, dtype=dt, compression="lzf")[...] = val
            elif key in plain:
                fh.create_dataset(key, val.shape, dtype=dt)[...] = val
            else:
                fh.create_dataset(key, val.shape, dtype=dt, compression="gzip", compression_opts=4)[...] = val
    return filename


def restart_mcmc(PSystem, record_file, in_path="./", out_path="./", itmax=100, intra_steps=2, suffix="_2",
                 restart_ladder=False):
    """``MCMC.restart_mcmc`` (mcmc.py:404-500) for the records ``run_mcmc`` writes: the keyword
    arguments that continue a run from its last saved ensemble -- ``run_mcmc(PSystem, **args)``.
    The ladder is the last saved one, or the first (``restart_ladder=True``); the new file is the
    old name plus ``suffix``. ``record_file``: the ``.npz`` (or, with h5py, ``.hdf5``) file name."""
    import os
    if not os.path.exists(out_path):
        raise RuntimeError(f"directory {out_path} does not exists")
    if not out_path.endswith("/"):
        out_path += "/"
    if record_file == "":
        raise RuntimeError("To restart, provide 'PSystem' and 'hdf5_file'")
    assert suffix != "", "New HDF5 file name cannot coincide with previous run. Try changing 'suffix' name."
    file_in = (in_path if in_path.endswith("/") else in_path + "/") + record_file
    if record_file.endswith(".npz"):
        z = np.load(file_in)
        index, allbetas, chains = int(z["INDEX"][0]), z["BETAS"], z["CHAINS"]
    else:
        import h5py
        with h5py.File(file_in, "r") as fh:
            index, allbetas, chains = int(fh["INDEX"][0]), fh["BETAS"][...], fh["CHAINS"][...]
    ladder = allbetas[0] if restart_ladder else allbetas[index]
    print("\n =========== RESTARTING MCMC ===========\n")
    print("--> Restarting from file: ", file_in)
    print("--> Temperature ladder status: ", "Restarting temperature ladder" if restart_ladder
          else "Temperature ladder continued")
    print()
    return {"p0": np.array(chains[:, :, index, :]), "betas": np.array(ladder), "itmax": itmax,
            "intra_steps": intra_steps, "file_name": record_file.split(".")[-2], "path": out_path, "suffix": suffix}


def extract_best_solutions(PSystem, record, write_file=True, best_file=None, to_physical=None):
    """``utils.extract_best_solutions`` (utils.py:874-934) for a ``run_mcmc`` record (the dict, or
    the name of its ``.npz`` file): the saved best solution of every look, in physical form,
    without repeated MAP values, best first; list of ``(MAP, physical vector)``. With
    ``write_file`` the ``.best`` text table of the reference (header with Mstar, Rstar, Nplanets,
    the parameter names, then one line per solution).
    to_physical: X_cube[B, ndim] -> flat[B, 7*npla]; default the GPU engine of PSystem."""
    from .optimizers import _writefile
    name = None
    if not isinstance(record, dict):
        name = str(record)
        record = np.load(name)
    index = int(record["INDEX"][0])
    cubes = np.asarray(record["BESTSOLS"][:index + 1], dtype=float)
    if to_physical is None:
        flat = _u.engine_for(PSystem).cube_to_physical(cubes)[0]
    else:
        flat = np.asarray(to_physical(cubes))
    pairs = list(dict(zip([float(v) for v in record["MAP"][:index + 1]], list(flat))).items())
    ranked = sorted(pairs, key=lambda t: t[0], reverse=True)
    if write_file:
        if best_file is None:
            if name is None:
                raise ValueError("best_file is needed when the record is not a file")
            best_file = name.rsplit(".", 1)[0] + ".best"
        names = getattr(PSystem, "params_names_all", "") or ""
        _writefile(best_file, "w", "#Mstar[Msun]      Rstar[Rsun]     Nplanets", "%-10s " * 3 + "\n")
        _writefile(best_file, "a", "#{}            {}           {}\n".format(record["MSTAR"][0], record["RSTAR"][0],
                                                                             record["NPLA"][0]), "%-10s " * 3 + "\n")
        head = "#MAP   " + names
        _writefile(best_file, "a", head, "%-16s" + " %-11s" * (len(head.split()) - 1) + "\n")
        for m, sol in ranked:
            text = " " + str(m) + " " + " ".join(str(v) for v in sol)
            _writefile(best_file, "a", text, "%-30s" + " %-11s" * (len(text.split()) - 1) + "\n")
        print(f"--> Best solutions from the run will be written at: {best_file}")
    return ranked
