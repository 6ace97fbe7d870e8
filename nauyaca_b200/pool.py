"""Batch seams that already exist in the reference's callers.

GpuPool         : drop-in for the ``pool=`` object handed to ptemcee.Sampler (mcmc.py:193-211).
                  ptemcee only calls ``pool.map(evaluator, list_of_vectors)``; here the whole
                  list goes to the GPU as one batch.
vectorized_chi2 : objective for scipy.optimize.differential_evolution(vectorized=True,
                  updating='deferred') (optimizers.py:78-83): x of shape (ndim, S) -> (S,).
"""
import numpy as np

from . import utils as _u


def _is_our_logl(f):
    return getattr(f, "__name__", "") == "log_likelihood_func"


def _is_our_chi2(f):
    return getattr(f, "__name__", "") == "calculate_chi2"


class GpuPool:
    """``map(func, iterable)`` that recognises the likelihood callables of the hot path.

    * ptemcee's LikePriorEvaluator (attributes logl, logp, loglargs, logpargs, loglkwargs,
      logpkwargs): the prior is evaluated on the host per vector (MCMC.logprior is 0.0 by
      default, mcmc.py:505-509, but users assign arbitrary Python), the likelihood of every
      vector with a finite prior in ONE GPU batch. Returns [(logl, logp), ...] exactly as the
      evaluator would (logl = 0 where logp = -inf; NaN raises ValueError).
    * functools.partial / plain log_likelihood_func or calculate_chi2 with the PSystem bound
      through ``args``: one GPU batch.
    * anything else is not part of the accelerated path and is mapped serially.
    """

    def __init__(self, PSystem=None, device=None):
        self.PSystem = PSystem
        self.device = device

    # multiprocessing.Pool API subset used by the reference (mcmc.py:193, closing(...))
    def close(self):
        pass

    def join(self):
        pass

    def terminate(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def map(self, func, iterable, chunksize=None):
        xs = [np.asarray(x, dtype=np.float64) for x in iterable]
        if not xs:
            return []
        logl = getattr(func, "logl", None)
        logp = getattr(func, "logp", None)
        if logl is not None and logp is not None and _is_our_logl(logl):
            largs = list(getattr(func, "loglargs", []) or [])
            lkw = dict(getattr(func, "loglkwargs", {}) or {})
            pargs = list(getattr(func, "logpargs", []) or [])
            pkw = dict(getattr(func, "logpkwargs", {}) or {})
            PS = largs[0] if largs else self.PSystem
            cube = largs[1] if len(largs) > 1 else lkw.get("cube", True)
            ins = largs[2] if len(largs) > 2 else lkw.get("insert_constants", False)
            lps = [logp(x, *pargs, **pkw) for x in xs]
            for lp in lps:
                if np.isnan(lp):
                    raise ValueError('Prior function returned NaN.')
            idx = [i for i, lp in enumerate(lps) if lp != float('-inf')]
            lls = np.zeros(len(xs))
            if idx:
                vals = _u.log_likelihood_batch(np.stack([xs[i] for i in idx]), PS, cube=cube, insert_constants=ins)
                if np.isnan(vals).any():
                    raise ValueError('Log likelihood function returned NaN.')
                lls[idx] = vals
            return [(float(ll) if lp != float('-inf') else 0, lp) for ll, lp in zip(lls, lps)]
        base = getattr(func, "func", func)          # functools.partial
        pargs = list(getattr(func, "args", ()) or ())
        pkw = dict(getattr(func, "keywords", {}) or {})
        if _is_our_logl(base):
            PS = pargs[0] if pargs else pkw.get("PSystem", self.PSystem)
            cube = pargs[1] if len(pargs) > 1 else pkw.get("cube", True)
            ins = pargs[2] if len(pargs) > 2 else pkw.get("insert_constants", False)
            return [float(v) for v in _u.log_likelihood_batch(np.stack(xs), PS, cube=cube, insert_constants=ins)]
        if _is_our_chi2(base):
            PS = pargs[0] if pargs else pkw.get("PSystem", self.PSystem)
            return [float(v) for v in _u.calculate_chi2_batch(np.stack(xs), PS)]
        return [func(x) for x in xs]


def vectorized_chi2(PSystem):
    """Objective for scipy's vectorized differential evolution: f(x[ndim, S]) -> chi2[S].
    Every member of a DE generation is evaluated in one GPU batch."""
    def f(x, *_args):
        x = np.asarray(x, dtype=np.float64)
        if x.ndim == 1:
            return float(_u.calculate_chi2_batch(x[None, :], PSystem)[0])
        return _u.calculate_chi2_batch(np.ascontiguousarray(x.T), PSystem)
    return f
