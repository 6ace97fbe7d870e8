"""CPU: the parallel-tempering sampler (nauyaca_b200/ptsampler.py) — the ptemcee surface that
mcmc.py:195-284 touches, statistical correctness on a known target, batched evaluation."""
import numpy as np
import pytest

from nauyaca_b200.ptsampler import Sampler, default_beta_ladder, LikePriorEvaluator


def _gauss_logl(x, mu, icov):
    d = np.asarray(x) - mu
    return -0.5 * float(d @ icov @ d)


def _flat_prior(x):
    return 0.0 if np.all(np.abs(x) < 50) else -np.inf


def test_beta_ladder():
    b = default_beta_ladder(11, ntemps=16, Tmax=100.0)
    assert len(b) == 16 and b[0] == 1.0 and abs(b[-1] - 0.01) < 1e-15 and np.all(np.diff(b) < 0)
    b = default_beta_ladder(18, ntemps=5)                    # no Tmax: hottest chain at beta = 0
    assert len(b) == 5 and b[0] == 1.0 and b[-1] == 0.0
    with pytest.raises(ValueError):
        default_beta_ladder(3)
    with pytest.raises(ValueError):
        default_beta_ladder(3, ntemps=4, Tmax=0.5)


def test_constructor_checks():
    with pytest.raises(ValueError):
        Sampler(7, 3, _gauss_logl, _flat_prior, ntemps=2, Tmax=10)          # odd
    with pytest.raises(ValueError):
        Sampler(4, 3, _gauss_logl, _flat_prior, ntemps=2, Tmax=10)          # < 2*dim
    s = Sampler(8, 3, _gauss_logl, _flat_prior, ntemps=2, Tmax=10, loglargs=[np.zeros(3), np.eye(3)])
    with pytest.raises(ValueError):
        list(s.sample(np.full((2, 8, 3), 100.0), iterations=1))             # outside prior support


def test_gaussian_target_and_surface():
    rng = np.random.mtrand.RandomState(12345)
    dim, nw, nt = 3, 40, 4
    mu = np.array([1.0, -2.0, 0.5])
    cov = np.array([[1.0, 0.6, 0.0], [0.6, 2.0, 0.3], [0.0, 0.3, 0.5]])
    icov = np.linalg.inv(cov)
    s = Sampler(nw, dim, _gauss_logl, _flat_prior, ntemps=nt, Tmax=30.0, a=2.0, loglargs=[mu, icov],
                adaptation_lag=1000, adaptation_time=100, random=rng)
    p0 = rng.normal(size=(nt, nw, dim))
    n = 0
    for p, logpost, logl in s.sample(p0, iterations=1500, thin=5, storechain=True, adapt=True):
        n += 1
        assert p.shape == (nt, nw, dim) and logpost.shape == (nt, nw) and logl.shape == (nt, nw)
    assert n == 1500 and s.chain.shape == (nt, nw, 300, dim)
    cold = s.chain[0, :, 60:, :].reshape(-1, dim)
    assert np.allclose(cold.mean(axis=0), mu, atol=0.12)
    assert np.allclose(np.cov(cold.T), cov, atol=0.25)
    hot = s.chain[-1, :, 60:, :].reshape(-1, dim)
    assert np.all(np.var(hot, axis=0) > 3 * np.var(cold, axis=0))          # tempered chain is wider
    # surface used by mcmc.py:240-284
    assert s.get_autocorr_time().shape == (nt, dim)
    assert s.tswap_acceptance_fraction.shape == (nt,) and np.all(s.tswap_acceptance_fraction > 0)
    acc = s.acceptance_fraction
    assert acc.shape == (nt, nw) and 0.1 < acc[0].mean() < 0.9
    b = s.betas
    assert b[0] == 1.0 and abs(b[-1] - 1 / 30.0) < 1e-12 and np.all(np.diff(b) < 0)   # ends fixed, monotone
    # logpost = beta*logl + logp is maintained through swaps and ladder moves
    assert np.allclose(logpost, logl * b[:, None])


class _ListPool:
    def __init__(self):
        self.calls = []

    def map(self, f, it):
        xs = list(it)
        self.calls.append(len(xs))
        return [f(x) for x in xs]


def test_pool_and_batch_paths_are_identical():
    dim, nw, nt = 2, 12, 3
    mu, icov = np.zeros(dim), np.eye(dim)
    p0 = np.random.mtrand.RandomState(1).normal(size=(nt, nw, dim))
    pool = _ListPool()
    s1 = Sampler(nw, dim, _gauss_logl, _flat_prior, ntemps=nt, Tmax=10.0, a=10, loglargs=[mu, icov], pool=pool,
                 random=np.random.mtrand.RandomState(7))
    r1 = s1.run_mcmc(p0, iterations=25, adapt=True)
    # every half-step is ONE map over ntemps*nwalkers/2 vectors (plus the initial evaluation)
    assert pool.calls[0] == nt * nw and set(pool.calls[1:]) == {nt * nw // 2} and len(pool.calls) == 51
    batch = lambda X: -0.5 * np.einsum('bi,ij,bj->b', X - mu, icov, X - mu)
    s2 = Sampler(nw, dim, _gauss_logl, _flat_prior, ntemps=nt, Tmax=10.0, a=10, loglargs=[mu, icov],
                 batch_logl=batch, random=np.random.mtrand.RandomState(7))
    r2 = s2.run_mcmc(p0, iterations=25, adapt=True)
    assert np.allclose(r1[0], r2[0], atol=1e-12) and np.allclose(r1[2], r2[2], atol=1e-12)
    assert np.array_equal(s1.nprop_accepted, s2.nprop_accepted)


def test_evaluator_semantics():
    ev = LikePriorEvaluator(lambda x: float("nan"), lambda x: 0.0)
    with pytest.raises(ValueError):
        ev(np.zeros(2))
    ev = LikePriorEvaluator(lambda x: 1.0, lambda x: -np.inf)
    assert ev(np.zeros(2)) == (0, -np.inf)
