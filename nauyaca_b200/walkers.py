"""Initial walker populations for the parallel-tempering sampler: the host-side mirror of the
reference's ``init_walkers`` (nauyaca/utils.py:463-716).

Same signature, same four distributions ("Uniform", "Gaussian", "Picked", "Ladder"), same
errors, and the SAME sequence of draws from numpy's global generator, so that for a given
``np.random.seed`` the population equals the reference's bit for bit
(tests/golden/init_walkers.json, produced by importing the reference). The populations are
what ``MCMC(p0=...)`` and the samplers here take as ``p0``; on the GPU their likelihoods are one
batch call (``utils.log_likelihood_batch(p0.reshape(-1, ndim), PSystem)``), which is what the
10^7-point sweep of BASELINE config 5 does.
"""
import numpy as np

__all__ = ['init_walkers']

_SCATTER = 0.01      # width of the perturbation around an optimizer solution (utils.py:593, 691)


def _truncated_normal(center, lo=0.0, hi=1.0):
    """One draw of N(center, 0.01) repeated until it falls inside [lo, hi] (utils.py:592-598)."""
    while True:
        v = np.random.normal(center, _SCATTER)
        if lo <= v <= hi:
            return v


def _uniform(PSystem, ntemps, nwalkers):
    """utils.py:529-543: one vector of ntemps*nwalkers uniform draws per dimension."""
    print("\n--> Selected distribution: Uniform")
    n = ntemps * nwalkers
    cols = [np.random.uniform(lo, hi, n) for lo, hi in PSystem.hypercube]
    return np.array(cols).T.reshape(ntemps, nwalkers, len(PSystem.hypercube))


def _solutions(opt_data, fbest):
    """Optimizer results -> the best fraction of valid solutions, ascending chi2
    (utils.py:551-568). Returns (rows of [chi2, cube...], number of valid solutions)."""
    if type(opt_data) == dict:
        opt_data = np.column_stack((opt_data['chi2'], opt_data['cube']))
    opt_data = np.asarray(opt_data)
    cube = opt_data[:, 1:]
    assert (cube >= 0.).all() and (cube <= 1.0).all(), "Invalid opt_data. Provide 'cube' solutions"
    valid = opt_data[opt_data[:, 0] < 1e+20]          # 1e20: the optimizer found no solution
    ranked = sorted(valid, key=lambda row: row[0])
    return ranked[:int(len(ranked) * fbest)], len(ranked)


def _split(seq, n):
    """n sequential chunks of seq, the first len(seq) % n one longer (utils.py:710-715)."""
    d, r = divmod(len(seq), n)
    out, start = [], 0
    for i in range(n):
        stop = start + d + (1 if i < r else 0)
        out.append(seq[start:stop])
        start = stop
    return out


def _from_opt(PSystem, distribution, ntemps, nwalkers, opt_data, fbest):
    kind = distribution.lower()
    best, n_valid = _solutions(opt_data, fbest)
    params = [row[1:] for row in best]
    print(f"    {len(best)} of {n_valid} solutions taken")
    ndim_b, ndim_c = len(PSystem.bounds), len(PSystem.hypercube)
    pop = []
    if kind == 'picked':
        # a random solution per walker, every coordinate scattered inside [0, 1]
        for _ in range(ntemps * nwalkers):
            sol = params[np.random.choice(range(len(params)))]
            pop.append([_truncated_normal(v) for v in sol.tolist()])
        return np.array(pop).reshape(ntemps, nwalkers, ndim_b)
    if kind == 'gaussian':
        # per dimension: N(mean, std) of the solutions, rejected outside the hypercube
        for d, col in enumerate(np.array(params).T):
            lo, hi = PSystem.hypercube[d][0], PSystem.hypercube[d][1]
            draws = []
            while len(draws) < ntemps * nwalkers:
                mu, sig = np.mean(col), np.std(col)
                v = mu if sig == 0.0 else np.random.normal(loc=mu, scale=sig)
                if lo <= v and v <= hi:
                    draws.append(v)
            pop.append(draws)
        return np.array(pop).T.reshape(ntemps, nwalkers, ndim_b)
    if kind == 'ladder':
        # temperature t draws from the best (t+1)/ntemps of the solutions
        chunks = _split(params, ntemps)
        for t in range(ntemps):
            pool = [row for c in chunks[:t + 1] for row in c]
            for i in np.random.choice(range(len(pool)), nwalkers):
                pop.append([_truncated_normal(v) for v in pool[i]])
        return np.array(pop).reshape(ntemps, nwalkers, ndim_c)
    return pop


def init_walkers(PSystem, distribution=None, opt_data=None, ntemps=None, nwalkers=None, fbest=1.0):
    """Initial population of walkers, shape (ntemps, nwalkers, ndim), in hypercube coordinates.

    PSystem needs ``ndim``, ``hypercube`` and ``bounds``. ``opt_data`` is the optimizers' result
    dictionary (keys 'chi2', 'cube') or the array of a ``*_cube.opt`` file (chi2 in column 0).
    """
    if nwalkers < 2 * PSystem.ndim:
        raise RuntimeError(f"Number of walkers must be >= 2*ndim, i.e., " +
                           f"nwalkers have to be >= {2*PSystem.ndim}.")
    if distribution.lower() == 'uniform':
        return _uniform(PSystem, ntemps, nwalkers)
    if opt_data is None:
        print("Arguments not understood")
        return None
    assert (0.0 < fbest <= 1.0), "fbest must be between 0 and 1"
    if distribution.lower() in ("gaussian", "picked", "ladder"):
        print("\n--> Selected distribution: {}".format(distribution))
        return _from_opt(PSystem, distribution, ntemps, nwalkers, opt_data, fbest)
    raise ValueError("--> Argument 'distribution' does not match with any supported distribution. "
                     "Available options are: \n *Uniform \n *Gaussian \n *Picked \n *Ladder "
                     "\n For the last three, please provide results from the optimizer routine "
                     "through 'opt_data' argument.")
