"""CPU: host-side logic of the package (no GPU compute): reference-interface helpers,
the ptemcee pool seam, the DE seam, multi-rank sharding over gloo."""
import math
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT


def test_shard_bounds_cover_and_balance():
    from nauyaca_b200.dist import shard_bounds
    for B in (0, 1, 7, 8, 5000, 100000, 10 ** 7 + 3):
        for W in (1, 2, 4, 8):
            cuts = [shard_bounds(B, W, r) for r in range(W)]
            assert cuts[0][0] == 0 and cuts[-1][1] == B
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(W - 1))
            sizes = [h - l for l, h in cuts]
            assert max(sizes) - min(sizes) <= 1


def test_sharded_evaluator_gloo(oracle, systems, tmp_path):
    """world_size 2 over gloo: every rank evaluates only its shard and ends with the full
    result, identical to the single-process evaluation."""
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    outs = [str(tmp_path / f"r{r}.npz") for r in range(2)]
    procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_gloo_worker.py"), str(r), "2",
                               str(port), outs[r]]) for r in range(2)]
    for p in procs:
        assert p.wait(timeout=240) == 0
    spec = systems["doc2"]
    X = np.random.default_rng(77).random((37, spec["ndim"])); X[5, 0] = 2.0
    chi2, logl, st = oracle.OracleSystem(spec).loglike(X, mode=0)
    for o in outs:
        d = np.load(o)
        assert np.array_equal(d["chi2"], chi2) and np.array_equal(d["logl"], logl) and np.array_equal(d["st"], st)
    assert chi2[5] == math.inf


def test_reference_helpers(systems, plumbing, spec_ps):
    from nauyaca_b200 import utils as U
    PS = spec_ps(systems["sys3"])
    # intervals: list returned at the first False (utils.py:1193-1219)
    assert U.intervals(PS.hypercube, [0.5] * 18) == [True] * 18
    assert U.intervals(PS.hypercube, [0.5, 1.5, 0.5]) == [True, False]
    assert U.intervals(PS.hypercube, [0.0, 1.0]) == [True, True]
    assert U.intervals(PS.hypercube, [float("nan")]) == [False]
    # _insert_constants / _remove_constants round trip (utils.py:440-460)
    x = list(np.arange(18, dtype=float))
    full = U._insert_constants(PS, x)
    assert len(full) == 21 and full[3] == 90.0 and full[17] == 90.3 and full[20] == 88.5
    assert list(U._remove_constants(PS, full)) == x
    # _chi2 with a missing epoch and a missing planet entry (utils.py:161-216)
    obs = {"b": {0: 1.0, 1: 2.0}, "c": {0: 5.0}}
    sig = {"b": {0: 0.5, 1: 0.5}, "c": {0: 1.0}}
    sim = {"b": {0: 1.5}, "c": {0: 5.0}}
    assert U._chi2(obs, sig, sim) == 1.0 + 1e20
    assert U._chi2(obs, sig, sim, individual=True) == {"b": 1.0 + 1e20, "c": 0.0}
    # _ephemeris masks rsky > rstarAU and keeps epoch gaps (utils.py:84-123)
    SP = np.array([[0, 0, 1, 0], [0, 1, 0, 2], [1.0, 2.0, 1.5, 3.0], [0.0, 1.0, 0.0, 0.0], [0, 0, 0, 0.0]])
    e = U._ephemeris(PS, SP)
    assert e["Planet-b"] == {0: 1.0, 2: 3.0} and e["Planet-c"] == {0: 1.5} and e["Planet-d"] == {}


def test_run_ttvfast_rejects_bad_length():
    from nauyaca_b200 import utils as U
    with pytest.raises(SystemExit):
        U.run_TTVFast([1.0] * 13, 1.0, 0.0, 10.0, 0.1)       # utils.py:48-50


class _Evaluator:
    """Shape of ptemcee's LikePriorEvaluator."""

    def __init__(self, logl, logp, loglargs):
        self.logl, self.logp, self.loglargs, self.logpargs = logl, logp, loglargs, []
        self.loglkwargs, self.logpkwargs = {}, {}


def test_gpupool_protocol(monkeypatch, oracle, systems, spec_ps):
    """GpuPool.map on a ptemcee-style evaluator: one batched likelihood call, host prior,
    (0, -inf) where the prior is -inf, NaN raises. The batch call is stubbed with the oracle
    here (CPU); tests/test_gpu_api.py runs the same through the GPU."""
    import nauyaca_b200 as nb
    from nauyaca_b200 import utils as U
    spec = systems["doc2"]; PS = spec_ps(spec); osys = oracle.OracleSystem(spec)
    calls = []

    def fake_batch(X, PSystem, cube=True, insert_constants=False):
        calls.append(np.asarray(X).shape)
        return osys.loglike(X, mode=0 if cube else 1)[1]
    monkeypatch.setattr(U, "log_likelihood_batch", fake_batch)
    rng = np.random.default_rng(1)
    xs = list(rng.random((10, spec["ndim"])))
    prior = lambda x, **kw: -np.inf if x[0] > 0.8 else 0.0
    ev = _Evaluator(U.log_likelihood_func, prior, [PS, True, False])
    res = nb.GpuPool().map(ev, xs)
    n_ok = sum(1 for x in xs if x[0] <= 0.8)
    assert calls == [(n_ok, spec["ndim"])]
    ref = osys.loglike(np.stack(xs), mode=0)[1]
    for x, (ll, lp), r in zip(xs, res, ref):
        if x[0] > 0.8:
            assert ll == 0 and lp == -np.inf
        else:
            assert ll == r and lp == 0.0
    ev_nan = _Evaluator(U.log_likelihood_func, lambda x: float("nan"), [PS, True, False])
    with pytest.raises(ValueError):
        nb.GpuPool().map(ev_nan, xs)
    # unknown callables are not part of the accelerated path: plain serial map
    assert nb.GpuPool().map(lambda x: float(x.sum()), xs[:3]) == [float(x.sum()) for x in xs[:3]]


def test_vectorized_chi2_shapes(monkeypatch, oracle, systems, spec_ps):
    import nauyaca_b200 as nb
    from nauyaca_b200 import utils as U
    spec = systems["doc2"]; PS = spec_ps(spec); osys = oracle.OracleSystem(spec)
    monkeypatch.setattr(U, "calculate_chi2_batch", lambda X, P: osys.loglike(X, mode=0)[0])
    f = nb.vectorized_chi2(PS)
    X = np.random.default_rng(3).random((spec["ndim"], 25))       # scipy passes (ndim, S)
    out = f(X)
    assert out.shape == (25,)
    assert np.array_equal(out, osys.loglike(np.ascontiguousarray(X.T), mode=0)[0])
    assert f(X[:, 0]) == out[0]


def test_sharded_mcmc_gloo(oracle, systems, spec_ps, tmp_path):
    """The multi-rank MCMC path: replicated sampler state, every likelihood block sharded and
    all-gathered. Both ranks must hold the chain of the single-process run, and each must have
    evaluated only half of every block."""
    import nauyaca_b200 as nb
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    outs = [str(tmp_path / f"m{r}.npz") for r in range(2)]
    procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_gloo_mcmc_worker.py"), str(r), "2",
                               str(port), outs[r]]) for r in range(2)]
    for p in procs:
        assert p.wait(timeout=300) == 0
    spec = systems["doc2"]; PS = spec_ps(spec); osys = oracle.OracleSystem(spec)
    rng = np.random.mtrand.RandomState(21)
    p0 = np.clip(0.5 + 0.05 * rng.normal(size=(3, 24, spec["ndim"])), 0, 1)
    ref = nb.mcmc.make_sampler(PS, p0, tmax=20.0, random=np.random.mtrand.RandomState(4),
                               evaluate=lambda X: osys.loglike(X, mode=0)[1])
    p, logpost, logl = ref.run_mcmc(p0, iterations=6, adapt=True)
    for o in outs:
        d = np.load(o)
        assert np.array_equal(d["p"], p) and np.array_equal(d["logl"], logl)
        assert d["nloc"][0] == 36 and np.all(d["nloc"][1:] == 18)       # 72 initial / 36 per half-step, halved
