"""Parallel-tempered affine-invariant ensemble sampler with a BATCHED likelihood.

nauyaca's MCMC driver (reference mcmc.py:195-221) uses ``ptemcee.Sampler`` (ptemcee 1.0.0,
Vousden, Farr & Mandel 2016), which is neither vendored in the reference nor installable
here. This module restates that sampler's published algorithm behind the same constructor /
``sample()`` / attribute surface that mcmc.py touches, so that ``MCMC.run`` can be pointed at
it unchanged:

    Sampler(nwalkers, dim, logl, logp, ntemps=, betas=, Tmax=, a=, pool=, loglargs=,
            logpkwargs=, adaptation_lag=, adaptation_time=)
    sampler.sample(p0, iterations=, thin=, storechain=, adapt=, swap_ratios=)  -> generator
    sampler.chain, .betas, .acceptance_fraction, .tswap_acceptance_fraction,
    sampler.get_autocorr_time()

Per iteration: two half-ensemble stretch moves with z = exp(U(-ln a,
ln a)); accept with dim*ln z + beta*dlogl + dlogp; temperature swaps between adjacent
ladder rungs from the hottest down, on random walker permutations; optional Vousden ladder
adaptation with decay lag/(t+lag)/time. Each half-step needs the likelihood of
ntemps*nwalkers/2 proposals: they go to the GPU as ONE batch, either through
``pool.map`` exactly as ptemcee does (``pool=GpuPool()``) or, faster, through
``batch_logl(X[B, dim]) -> logl[B]`` which skips the per-vector Python objects (and
``batch_logp`` for a vectorised prior; the per-vector prior loop costs more than the GPU
kernel at 5000 proposals per half-step).

Parity note: ptemcee draws from an unseeded numpy RandomState, so its chains are not
reproducible either; parity with the reference is per likelihood evaluation (the GPU path),
and statistical for the sampler (tests/test_ptsampler.py).
"""
import numpy as np

__all__ = ["Sampler", "DeviceSampler", "default_beta_ladder", "LikePriorEvaluator"]


def default_beta_ladder(ndim, ntemps=None, Tmax=None):
    """Geometric ladder of inverse temperatures. With ntemps and Tmax (the case mcmc.py:195-205
    produces when ``tmax`` is set) betas = logspace(0, -log10(Tmax), ntemps). Without Tmax the
    ratio between rungs is the dimension-dependent spacing that gives ~25 % swap acceptance on
    a Gaussian, 1 + 2 sqrt(ln 4)/sqrt(ndim) (Vousden et al. 2016, eq. 22 asymptote) and, with
    ntemps given, the hottest chain is put at beta = 0."""
    if type(ndim) != int or ndim < 1:
        raise ValueError('Invalid number of dimensions specified.')
    if ntemps is None and Tmax is None:
        raise ValueError('Must specify one of ``ntemps`` and ``Tmax``.')
    if Tmax is not None and Tmax <= 1:
        raise ValueError('``Tmax`` must be greater than 1.')
    if ntemps is not None and (type(ntemps) != int or ntemps < 1):
        raise ValueError('Invalid number of temperatures specified.')
    tstep = 1.0 + 2.0 * np.sqrt(np.log(4.0)) / np.sqrt(ndim)
    appendInf = False
    if ntemps is None:
        ntemps = int(np.log(Tmax) / np.log(tstep) + 2)
    elif Tmax is None:
        appendInf = True
        ntemps -= 1
    if Tmax is not None and appendInf is False and ntemps > 1:
        tstep = np.exp(np.log(Tmax) / (ntemps - 1))
    betas = np.logspace(0, -np.log10(tstep) * (ntemps - 1), ntemps) if ntemps > 0 else np.array([])
    if appendInf:
        betas = np.concatenate((betas, [0]))
    return betas


class LikePriorEvaluator:
    """Picklable (logl, logp) evaluator of one vector: what ptemcee hands to ``pool.map``."""

    def __init__(self, logl, logp, loglargs=(), logpargs=(), loglkwargs=None, logpkwargs=None):
        self.logl, self.logp = logl, logp
        self.loglargs, self.logpargs = list(loglargs), list(logpargs)
        self.loglkwargs, self.logpkwargs = dict(loglkwargs or {}), dict(logpkwargs or {})

    def __call__(self, x):
        lp = self.logp(x, *self.logpargs, **self.logpkwargs)
        if np.isnan(lp):
            raise ValueError('Prior function returned NaN.')
        if lp == float('-inf'):
            return 0, lp           # cannot return -inf: it would break beta = 0
        ll = self.logl(x, *self.loglargs, **self.loglkwargs)
        if np.isnan(ll).any():
            raise ValueError('Log likelihood function returned NaN.')
        return ll, lp


class Sampler:
    def __init__(self, nwalkers, dim, logl, logp, ntemps=None, Tmax=None, betas=None, threads=1, pool=None,
                 a=2.0, loglargs=(), logpargs=(), loglkwargs=None, logpkwargs=None, adaptation_lag=10000,
                 adaptation_time=100, random=None, batch_logl=None, batch_logp=None):
        self._random = np.random.mtrand.RandomState() if random is None else random
        self._likeprior = LikePriorEvaluator(logl, logp, loglargs, logpargs, loglkwargs, logpkwargs)
        self.a = a
        self.nwalkers = nwalkers
        self.dim = dim
        self.adaptation_time = adaptation_time
        self.adaptation_lag = adaptation_lag
        self._betas = (default_beta_ladder(self.dim, ntemps=ntemps, Tmax=Tmax) if betas is None
                       else np.array(betas, dtype=float).copy())
        self.ntemps = len(self._betas)
        if self.nwalkers % 2 != 0:
            raise ValueError('The number of walkers must be even.')
        if self.nwalkers < 2 * self.dim:
            raise ValueError('The number of walkers must be greater than ``2*dimension``.')
        self.pool = pool
        self.batch_logl = batch_logl
        self.batch_logp = batch_logp      # optional vectorised prior X[B, dim] -> logp[B]
        self.reset()

    # ------------------------------------------------------------------ state
    def reset(self, random=None, betas=None, time=None):
        if random is not None:
            self._random = random
        if betas is not None:
            self._betas = np.array(betas, dtype=float).copy()
        self._time = 0 if time is None else time
        self._p0 = None
        self._logposterior0 = None
        self._loglikelihood0 = None
        self.nswap = np.zeros(self.ntemps, dtype=float)
        self.nswap_accepted = np.zeros(self.ntemps, dtype=float)
        self.nprop = np.zeros((self.ntemps, self.nwalkers), dtype=float)
        self.nprop_accepted = np.zeros((self.ntemps, self.nwalkers), dtype=float)
        self._chain = None
        self._logposterior = None
        self._loglikelihood = None
        self._beta_history = None

    def run_mcmc(self, *args, **kwargs):
        res = None
        for res in self.sample(*args, **kwargs):
            pass
        return res

    # ------------------------------------------------------------------ evaluation
    def _evaluate(self, ps):
        """logl, logp of ps[ntemps, n, dim], each of shape (ntemps, n): ONE batch."""
        flat = np.ascontiguousarray(ps.reshape((-1, self.dim)))
        if self.batch_logl is not None:
            if self.batch_logp is not None:
                lp = np.asarray(self.batch_logp(flat), dtype=float)
            else:
                lp = np.array([self._likeprior.logp(x, *self._likeprior.logpargs, **self._likeprior.logpkwargs)
                               for x in flat], dtype=float)
            if np.isnan(lp).any():
                raise ValueError('Prior function returned NaN.')
            ok = lp != -np.inf
            ll = np.zeros(flat.shape[0])
            if ok.any():
                ll[ok] = np.asarray(self.batch_logl(flat[ok]), dtype=float)
            if np.isnan(ll).any():
                raise ValueError('Log likelihood function returned NaN.')
        else:
            mapf = map if self.pool is None else self.pool.map
            results = list(mapf(self._likeprior, flat))
            ll = np.fromiter((r[0] for r in results), float, count=len(results))
            lp = np.fromiter((r[1] for r in results), float, count=len(results))
        return ll.reshape((self.ntemps, -1)), lp.reshape((self.ntemps, -1))

    def _tempered_likelihood(self, logl, betas=None):
        betas = (self._betas if betas is None else betas).reshape((-1, 1))
        with np.errstate(invalid='ignore'):
            loglT = logl * betas
        loglT[np.isnan(loglT)] = -np.inf
        return loglT

    # ------------------------------------------------------------------ sampling
    def sample(self, p0=None, iterations=1, thin=1, storechain=True, adapt=False, swap_ratios=False):
        if p0 is not None:
            self._p0 = p = np.array(p0, dtype=float).copy()
            self._logposterior0 = None
            self._loglikelihood0 = None
        elif self._p0 is not None:
            p = self._p0
        else:
            raise ValueError('Initial walker positions not specified.')
        if p.shape != (self.ntemps, self.nwalkers, self.dim):
            raise ValueError('p0 must have shape (ntemps, nwalkers, dim).')
        if self._logposterior0 is None or self._loglikelihood0 is None:
            logl, logp = self._evaluate(p)
            logpost = self._tempered_likelihood(logl) + logp
            self._loglikelihood0, self._logposterior0 = logl, logpost
        else:
            logl, logpost = self._loglikelihood0, self._logposterior0
        if (logpost == -np.inf).any():
            raise ValueError('Attempting to start with samples outside posterior support.')
        if storechain:
            isave = self._expand_chain(iterations // thin)
        half = self.nwalkers // 2
        for _ in range(iterations):
            for j in (0, 1):
                jupdate, jsample = j, (j + 1) % 2
                pupdate = p[:, jupdate::2, :]
                psample = p[:, jsample::2, :]
                zs = np.exp(self._random.uniform(low=-np.log(self.a), high=np.log(self.a), size=(self.ntemps, half)))
                # partner of every updated walker, all temperatures at once
                js = self._random.randint(0, high=half, size=(self.ntemps, half))
                partner = np.take_along_axis(psample, js[:, :, None], axis=1)
                qs = partner + zs[:, :, None] * (pupdate - partner)
                qslogl, qslogp = self._evaluate(qs)
                qslogpost = self._tempered_likelihood(qslogl) + qslogp
                logpaccept = self.dim * np.log(zs) + qslogpost - logpost[:, jupdate::2]
                logr = np.log(self._random.uniform(low=0.0, high=1.0, size=(self.ntemps, half)))
                accepts = logr < logpaccept
                p[:, jupdate::2, :][accepts] = qs[accepts]
                logpost[:, jupdate::2][accepts] = qslogpost[accepts]
                logl[:, jupdate::2][accepts] = qslogl[accepts]
                self.nprop[:, jupdate::2] += 1.0
                self.nprop_accepted[:, jupdate::2] += accepts
            p, ratios = self._temperature_swaps(self._betas, p, logpost, logl)
            if adapt and self.ntemps > 1:
                dbetas = self._get_ladder_adjustment(self._time, self._betas, ratios)
                self._betas += dbetas
                logpost += self._tempered_likelihood(logl, betas=dbetas)
            if (self._time + 1) % thin == 0 and storechain:
                self._chain[:, :, isave, :] = p
                self._logposterior[:, :, isave] = logpost
                self._loglikelihood[:, :, isave] = logl
                self._beta_history[:, isave] = self._betas
                isave += 1
            self._time += 1
            self._p0, self._logposterior0, self._loglikelihood0 = p, logpost, logl
            if swap_ratios:
                yield p, logpost, logl, ratios
            else:
                yield p, logpost, logl

    def _temperature_swaps(self, betas, p, logpost, logl):
        ntemps = len(betas)
        ratios = np.zeros(ntemps - 1)
        for i in range(ntemps - 1, 0, -1):
            bi, bi1 = betas[i], betas[i - 1]
            dbeta = bi1 - bi
            iperm = self._random.permutation(self.nwalkers)
            i1perm = self._random.permutation(self.nwalkers)
            raccept = np.log(self._random.uniform(size=self.nwalkers))
            paccept = dbeta * (logl[i, iperm] - logl[i - 1, i1perm])
            self.nswap[i] += self.nwalkers
            self.nswap[i - 1] += self.nwalkers
            asel = paccept > raccept
            nacc = np.sum(asel)
            self.nswap_accepted[i] += nacc
            self.nswap_accepted[i - 1] += nacc
            ratios[i - 1] = nacc / self.nwalkers
            ia, i1a = iperm[asel], i1perm[asel]          # walkers of chain i / chain i-1 that trade places
            ptemp = p[i, ia, :]
            logltemp = logl[i, ia]
            logprtemp = logpost[i, ia]
            l1 = logl[i - 1, i1a]
            p[i, ia, :] = p[i - 1, i1a, :]
            logl[i, ia] = l1
            logpost[i, ia] = logpost[i - 1, i1a] - dbeta * l1
            p[i - 1, i1a, :] = ptemp
            logl[i - 1, i1a] = logltemp
            logpost[i - 1, i1a] = logprtemp + dbeta * logltemp
        return p, ratios

    def _get_ladder_adjustment(self, time, betas0, ratios):
        betas = betas0.copy()
        decay = self.adaptation_lag / (time + self.adaptation_lag)
        kappa = decay / self.adaptation_time
        dSs = kappa * (ratios[:-1] - ratios[1:])
        deltaTs = np.diff(1 / betas[:-1])
        deltaTs *= np.exp(dSs)
        betas[1:-1] = 1 / (np.cumsum(deltaTs) + 1 / betas[0])
        return betas - betas0

    def _expand_chain(self, nsave):
        if self._chain is None:
            isave = 0
            self._chain = np.zeros((self.ntemps, self.nwalkers, nsave, self.dim))
            self._logposterior = np.zeros((self.ntemps, self.nwalkers, nsave))
            self._loglikelihood = np.zeros((self.ntemps, self.nwalkers, nsave))
            self._beta_history = np.zeros((self.ntemps, nsave))
        else:
            isave = self._chain.shape[2]
            self._chain = np.concatenate((self._chain, np.zeros((self.ntemps, self.nwalkers, nsave, self.dim))), axis=2)
            self._logposterior = np.concatenate((self._logposterior, np.zeros((self.ntemps, self.nwalkers, nsave))), axis=2)
            self._loglikelihood = np.concatenate((self._loglikelihood, np.zeros((self.ntemps, self.nwalkers, nsave))), axis=2)
            self._beta_history = np.concatenate((self._beta_history, np.zeros((self.ntemps, nsave))), axis=1)
        return isave

    # ------------------------------------------------------------------ diagnostics
    def log_evidence_estimate(self, logls=None, fburnin=0.1):
        """Thermodynamic-integration estimate of ln Z and its error."""
        if logls is None:
            if self.loglikelihood is not None:
                logls = np.mean(self.loglikelihood, axis=1)[:, int(fburnin * self.loglikelihood.shape[2]):]
                logls = np.mean(logls, axis=1)
            else:
                raise ValueError('No log likelihood values available.')
        order = np.argsort(self._betas)[::-1]
        betas, logls = self._betas[order], logls[order]
        if betas[-1] != 0:
            betas = np.concatenate((betas, [0])); logls = np.concatenate((logls, [logls[-1]]))
        betas2 = np.concatenate((betas[::2], [0])) if betas[::2][-1] != 0 else betas[::2]
        logls2 = np.concatenate((logls[::2], [logls[-1]]))[:len(betas2)]
        logZ = -np.trapz(logls, betas)
        logZ2 = -np.trapz(logls2, betas2)
        return logZ, np.abs(logZ - logZ2)

    @property
    def random(self):
        return self._random

    @property
    def betas(self):
        return self._betas

    @property
    def time(self):
        return self._time

    @property
    def chain(self):
        return self._chain

    @property
    def flatchain(self):
        s = self.chain.shape
        return self._chain.reshape((s[0], -1, s[3]))

    @property
    def logprobability(self):
        return self._logposterior

    @property
    def loglikelihood(self):
        return self._loglikelihood

    @property
    def beta_history(self):
        return self._beta_history

    @property
    def tswap_acceptance_fraction(self):
        return self.nswap_accepted / self.nswap

    @property
    def acceptance_fraction(self):
        return self.nprop_accepted / self.nprop

    @property
    def acor(self):
        return self.get_autocorr_time()

    def get_autocorr_time(self, window=50):
        """Integrated autocorrelation time of the walker-averaged chain, (ntemps, dim)."""
        acors = np.zeros((self.ntemps, self.dim))
        for i in range(self.ntemps):
            x = np.mean(self._chain[i, :, :, :], axis=0)
            acors[i, :] = _autocorr_integrated_time(x, window=window)
        return acors


def _autocorr_function(x, axis=0):
    x = np.atleast_1d(x)
    n = x.shape[axis]
    f = np.fft.fft(x - np.mean(x, axis=axis, keepdims=True), n=2 * n, axis=axis)
    m = [slice(None)] * x.ndim
    m[axis] = slice(0, n)
    acf = np.fft.ifft(f * np.conjugate(f), axis=axis)[tuple(m)].real
    m[axis] = 0
    with np.errstate(invalid='ignore', divide='ignore'):
        return acf / acf[tuple(m)]


def _autocorr_integrated_time(x, axis=0, window=50):
    f = _autocorr_function(x, axis=axis)
    if f.ndim == 1:
        return 1 + 2 * np.sum(f[1:window])
    m = [slice(None)] * f.ndim
    m[axis] = slice(1, window)
    return 1 + 2 * np.sum(f[tuple(m)], axis=axis)


class DeviceSampler:
    """The same sampler with every move on the GPU (ttvb_pt_* in include/ttvb.h).

    Walkers, log-likelihoods and the ladder stay resident on the device; one PT iteration is a
    stream of small kernels around two likelihood launches, and the host only fetches the
    state when it is asked for. Restrictions: the reference's default flat prior
    (MCMC.logprior == 0.0, mcmc.py:505-509) and the TTV likelihood of a BatchEngine. Surface:
    the part of ptemcee.Sampler that mcmc.py:195-284 touches (sample(), chain, betas,
    acceptance_fraction, tswap_acceptance_fraction, get_autocorr_time())."""

    def __init__(self, engine, nwalkers, ntemps=None, Tmax=None, betas=None, a=2.0, adaptation_lag=10000,
                 adaptation_time=100, seed=None, cube=True, group=None, shard=False):
        """shard=True (with torch.distributed initialised, one process per GPU): the proposals of
        every half-step are split over the ranks of `group` and their log-likelihoods
        all-gathered (NCCL; host staging under gloo). Every rank must be given the same seed and
        the same p0, and then holds the same ensemble at every iteration."""
        import ctypes as C
        from . import _lib
        self._C, self._lib = C, _lib
        self.engine = engine
        self.dim = engine.ndim
        self.nwalkers = int(nwalkers)
        b = default_beta_ladder(self.dim, ntemps=ntemps, Tmax=Tmax) if betas is None else np.array(betas, dtype=float)
        self._betas = np.ascontiguousarray(b, dtype=np.float64)
        self.ntemps = len(self._betas)
        if self.nwalkers % 2 != 0:
            raise ValueError('The number of walkers must be even.')
        if self.nwalkers < 2 * self.dim:
            raise ValueError('The number of walkers must be greater than ``2*dimension``.')
        self._world, self._rank, self._group = 1, 0, group
        if shard:
            import torch.distributed as dist
            if dist.is_initialized():
                self._world, self._rank = dist.get_world_size(group), dist.get_rank(group)
        if seed is None:
            seed = int(np.random.SeedSequence().generate_state(2, dtype=np.uint32).view(np.uint64)[0])
            if self._world > 1:
                # every rank must make the same moves: rank 0 draws the seed for all of them
                box = [seed]
                dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0,
                                           group=group)
                seed = int(box[0])
        elif self._world > 1:
            self._same_on_all_ranks(int(seed), "seed")
        if self._world > 1 and self.ntemps * (self.nwalkers // 2) < 18944:
            import warnings
            warnings.warn(f"sharded DeviceSampler: a half-step of {self.ntemps * (self.nwalkers // 2)} proposals is "
                          "latency-bound on ANY number of GPUs (below one warp per scheduler, 18 944 systems on a B200, "
                          "a launch takes about one warp-latency): sharding it shortens the iteration only through the "
                          "fewer events every warp then handles (measured, 10 x 1000 walkers: 1.92 ms on 1 GPU, 1.85 ms "
                          "sharded over 2, 1.79 ms over 8, profiles/r2x_bench_c3*). For throughput run one replica of "
                          "the ensemble per GPU instead.", stacklevel=2)
        self._pt = C.c_void_p()
        _lib.check(_lib.lib().ttvb_pt_create(engine._h, self.ntemps, self.nwalkers, self._betas.ctypes.data, float(a),
                                             float(adaptation_lag), float(adaptation_time), C.c_uint64(seed),
                                             _lib.MODE_CUBE if cube else _lib.MODE_PHYSICAL, C.byref(self._pt)))
        self._time = 0
        self._started = False
        self._chain = self._logposterior = self._loglikelihood = self._beta_history = None
        shape = (self.ntemps, self.nwalkers)
        self._p = np.empty(shape + (self.dim,)); self._logl = np.empty(shape)
        self.nprop = np.zeros(shape); self.nprop_accepted = np.zeros(shape)
        self.nswap = np.zeros(self.ntemps); self.nswap_accepted = np.zeros(self.ntemps)

    def _same_on_all_ranks(self, value, what):
        """A sharded run is only right if every rank holds the same ensemble: compare a value (or a
        digest) across the ranks of the group and raise on all of them if it differs."""
        import torch.distributed as dist
        got = [None] * self._world
        dist.all_gather_object(got, value, group=self._group)
        if any(g != got[0] for g in got):
            raise ValueError(f'sharded DeviceSampler: {what} differs between ranks: {got}')

    def close(self):
        if getattr(self, "_pt", None) is not None and self._pt.value:
            self._lib.lib().ttvb_pt_destroy(self._pt)
            self._pt = self._C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _fetch(self):
        d = lambda a: a.ctypes.data
        self._lib.check(self._lib.lib().ttvb_pt_get(self._pt, d(self._p), d(self._logl), d(self._betas), d(self.nprop),
                                                    d(self.nprop_accepted), d(self.nswap), d(self.nswap_accepted), None))
        with np.errstate(invalid='ignore'):
            logpost = self._logl * self._betas[:, None]
        logpost[np.isnan(logpost)] = -np.inf
        return self._p.copy(), logpost, self._logl.copy()

    def sample(self, p0=None, iterations=1, thin=1, storechain=True, adapt=False, swap_ratios=False):
        L = self._lib.lib()
        if p0 is not None:
            p0 = np.ascontiguousarray(p0, dtype=np.float64)
            if p0.shape != (self.ntemps, self.nwalkers, self.dim):
                raise ValueError('p0 must have shape (ntemps, nwalkers, dim).')
            if self._world > 1:
                import hashlib
                self._same_on_all_ranks(hashlib.sha1(p0.tobytes()).hexdigest(), "p0")
            self._lib.check(L.ttvb_pt_set_walkers(self._pt, p0.ctypes.data, None))
            self._started = True
            p, logpost, logl = self._fetch()
            if (logl == -np.inf).any():
                raise ValueError('Attempting to start with samples outside posterior support.')
        elif not self._started:
            raise ValueError('Initial walker positions not specified.')
        if storechain:
            nsave = iterations // thin
            z = lambda *s: np.zeros(s)
            if self._chain is None:
                isave = 0
                self._chain = z(self.ntemps, self.nwalkers, nsave, self.dim)
                self._logposterior = z(self.ntemps, self.nwalkers, nsave); self._loglikelihood = z(self.ntemps, self.nwalkers, nsave)
                self._beta_history = z(self.ntemps, nsave)
            else:
                isave = self._chain.shape[2]
                self._chain = np.concatenate((self._chain, z(self.ntemps, self.nwalkers, nsave, self.dim)), axis=2)
                self._logposterior = np.concatenate((self._logposterior, z(self.ntemps, self.nwalkers, nsave)), axis=2)
                self._loglikelihood = np.concatenate((self._loglikelihood, z(self.ntemps, self.nwalkers, nsave)), axis=2)
                self._beta_history = np.concatenate((self._beta_history, z(self.ntemps, nsave)), axis=1)
        for _ in range(iterations):
            self._run(1, adapt)
            p, logpost, logl = self._fetch()
            if (self._time + 1) % thin == 0 and storechain:
                self._chain[:, :, isave, :] = p
                self._logposterior[:, :, isave] = logpost
                self._loglikelihood[:, :, isave] = logl
                self._beta_history[:, isave] = self._betas
                isave += 1
            self._time += 1
            yield p, logpost, logl

    def advance(self, iterations, adapt=True):
        """Run `iterations` PT iterations without fetching anything in between (the fast path:
        nauyaca only looks at the state every `intra_steps` iterations, mcmc.py:227-229)."""
        if not self._started:
            raise ValueError('Initial walker positions not specified.')
        self._run(int(iterations), adapt)
        self._time += int(iterations)
        return self._fetch()

    # ---- iteration drivers
    def _run(self, iterations, adapt):
        L, chk = self._lib.lib(), self._lib.check
        if self._world == 1:
            chk(L.ttvb_pt_run(self._pt, int(iterations), 1 if adapt else 0, None))
            return
        lo, hi = self._shard_range()
        for _ in range(int(iterations)):
            for half in (0, 1):
                chk(L.ttvb_pt_half_begin(self._pt, half, lo, hi, None))
                self._gather_proposal_logl()
                chk(L.ttvb_pt_half_end(self._pt, half, 1 if adapt else 0, None))

    def _shard_range(self):
        """Equal chunks of ceil(B / world) proposals (the last ones may be short or empty), so
        that the all-gather can run in place on the padded device buffer."""
        B = self.ntemps * (self.nwalkers // 2)
        self._chunk = -(-B // self._world)
        lo = min(B, self._rank * self._chunk)
        return lo, min(B, lo + self._chunk)

    def _gather_proposal_logl(self):
        import torch
        import torch.distributed as dist
        C = self._C
        if getattr(self, "_qbuf", None) is None:
            ptr, n, cap = C.c_void_p(), C.c_int64(), C.c_int64()
            self._lib.check(self._lib.lib().ttvb_pt_proposal_logl(self._pt, C.byref(ptr), C.byref(n), C.byref(cap)))
            need = self._chunk * self._world
            if need > cap.value:
                raise ValueError(f'world size {self._world} too large for the proposal buffer')

            class _Dev:
                __cuda_array_interface__ = {"shape": (need,), "typestr": "<f8", "data": (ptr.value, False), "version": 2}
            self._qbuf = torch.as_tensor(_Dev(), device=f"cuda:{self.engine.device}")
        q, c, r = self._qbuf, self._chunk, self._rank
        mine = q.narrow(0, r * c, c)
        if dist.get_backend(self._group) == "nccl":
            dist.all_gather_into_tensor(q, mine, group=self._group)          # in place, over NVLink
        else:                                                                # gloo (CPU-side tests): host staging
            out = torch.empty(c * self._world, dtype=torch.float64)
            dist.all_gather_into_tensor(out, mine.cpu(), group=self._group)
            q.copy_(out)

    run_mcmc = Sampler.run_mcmc
    betas = property(lambda self: self._betas)
    time = property(lambda self: self._time)
    chain = property(lambda self: self._chain)
    logprobability = property(lambda self: self._logposterior)
    loglikelihood = property(lambda self: self._loglikelihood)
    beta_history = property(lambda self: self._beta_history)
    tswap_acceptance_fraction = property(lambda self: self.nswap_accepted / self.nswap)
    acceptance_fraction = property(lambda self: self.nprop_accepted / self.nprop)
    get_autocorr_time = Sampler.get_autocorr_time
