"""The sampler seam of the reference's MCMC driver (mcmc.py:97-221) on the batched GPU path.

Only the part of ``MCMC.run`` that feeds the hot path is mirrored: the choice between cube /
physical walkers (mcmc.py:129-137), the ptemcee settings (mcmc.py:182-211: a = 10,
adaptation_lag = 1000, adaptation_time = 100, adapt = True) and the ``sampler.sample`` loop.
HDF5 bookkeeping, printing and restart stay with the reference (out of
scope)."""
import numpy as np

from . import utils as _u
from .ptsampler import Sampler

__all__ = ["make_sampler", "make_device_sampler", "pt_sample"]


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
