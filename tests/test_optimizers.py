"""CPU: the rendezvous batching of the optimizer chains (nauyaca_b200/optimizers.py)."""
import threading

import numpy as np
import pytest

from nauyaca_b200.optimizers import RendezvousBatcher, de_pw_nm


class _PS:
    def __init__(self, ndim):
        self.ndim = ndim
        self.hypercube = [[0.0, 1.0]] * ndim
        self.npla = 1


def _rosen_like(X):
    X = np.atleast_2d(X)
    c = np.linspace(0.3, 0.7, X.shape[1])
    return 1e4 * np.sum((X - c) ** 2, axis=1) + 50.0 * np.sum((X[:, 1:] - X[:, :-1] ** 2 - c[1:] + c[:-1] ** 2) ** 2, axis=1)


def test_rendezvous_batches_all_waiting_workers():
    calls = []

    def batch(X):
        calls.append(X.shape[0])
        return X.sum(axis=1)
    b = RendezvousBatcher(batch, 5, greedy=False)
    res = {}

    def work(i):
        try:
            for k in range(3 + i):                    # workers finish at different times
                res[(i, k)] = b.evaluate(np.full((1 + (i % 2), 4), float(i + k)))
        finally:
            b.worker_done()
    ts = [threading.Thread(target=work, args=(i,)) for i in range(5)]
    [t.start() for t in ts]
    b.serve()
    [t.join() for t in ts]
    assert calls[0] == sum(1 + (i % 2) for i in range(5))          # first round: everyone in one batch
    assert len(calls) == 7                                            # longest worker needs 7 rounds
    for (i, k), v in res.items():
        assert np.all(v == 4.0 * (i + k)) and v.shape == (1 + (i % 2),)


def test_rendezvous_propagates_errors():
    b = RendezvousBatcher(lambda X: 1 / 0, 2)
    errs = []

    def work():
        try:
            b.evaluate(np.zeros((1, 2)))
        except RuntimeError as e:
            errs.append(e)
        finally:
            b.worker_done()
    ts = [threading.Thread(target=work) for _ in range(2)]
    [t.start() for t in ts]
    with pytest.raises(ZeroDivisionError):
        b.serve()
    [t.join() for t in ts]
    assert len(errs) == 2


def test_chains_equal_serial_scipy_and_are_batched():
    """Same seeds, same objective values -> each chain takes exactly the path the serial scipy
    calls of optimizers.py:78-113 take; the batches carry one request per live chain."""
    from scipy.optimize import differential_evolution, minimize
    PS = _PS(4)
    nsols = 6
    res, info = de_pw_nm(PS, nsols, seed=123, evaluate=_rosen_like, method="scipy",
                         powell_opts={"maxfev": 400}, nm_opts={"maxfev": 600}, greedy=False)
    assert max(info["batches"]) == nsols * 30 * 4                     # a DE generation of every chain at once
    assert np.median(info["batches"]) >= 2
    # the greedy policy (default: a batch leaves as soon as the coordinator is free) groups the same
    # evaluations differently and returns the same chains
    res_g, info_g = de_pw_nm(PS, nsols, seed=123, evaluate=_rosen_like, method="scipy",
                             powell_opts={"maxfev": 400}, nm_opts={"maxfev": 600})
    assert sum(info_g["batches"]) == sum(info["batches"])
    assert all(a[0] == b[0] and a[1] == b[1] for a, b in zip(res, res_g))
    ss = np.random.SeedSequence(123)
    seeds = [int(s.generate_state(1)[0]) for s in ss.spawn(nsols)]
    for i in range(nsols):
        box = {"lo": np.zeros(4), "hi": np.ones(4)}

        def f(x):
            x = np.asarray(x)
            X = x[None, :] if x.ndim == 1 else x.T
            v = np.where(np.all((X >= box["lo"]) & (X <= box["hi"]), axis=1), _rosen_like(X), np.inf)
            return float(v[0]) if x.ndim == 1 else v
        de = differential_evolution(f, [(0, 1)] * 4, seed=seeds[i], vectorized=True, updating="deferred", disp=False,
                                    popsize=30, maxiter=1, polish=False, mutation=(1.2, 1.7), recombination=0.6)
        box["lo"], box["hi"] = np.maximum(0, de.x - .3), np.minimum(1, de.x + .3)
        po = minimize(f, list(de.x), method="Powell", options=dict(maxiter=12000, maxfev=400, xtol=0.001, ftol=1, disp=False))
        nm = minimize(f, list(po.x), method="Nelder-Mead",
                      options=dict(maxiter=15000, maxfev=600, xatol=0.0001, fatol=1, disp=False, adaptive=True))
        assert info["stages"][i] == (de.fun, po.fun, nm.fun)
        assert res[i][0] == nm.fun and res[i][1] == list(nm.x)


def test_lockstep_local_search_matches_scipy_in_quality():
    """Powell and Nelder-Mead of all chains advanced together (lockstep.py): every move is one batch over
    the live chains, the stopping rules are scipy's, and the minima are as good as those of the serial
    scipy chains on the same objective."""
    from scipy.optimize import minimize
    from nauyaca_b200 import lockstep
    n, C = 5, 12
    rng = np.random.default_rng(7)
    x0 = 0.5 + 0.15 * rng.standard_normal((C, n))
    lo, hi = np.maximum(0, x0 - .3), np.minimum(1, x0 + .3)
    calls = []

    def ev(X):
        calls.append(X.shape[0])
        return _rosen_like(X)
    po = lockstep.powell(ev, x0, lo, hi)
    assert np.all(po["fun"] <= _rosen_like(x0)) and np.all((po["x"] >= lo) & (po["x"] <= hi))
    assert max(calls) >= C * 8                                   # a comb of all chains in one batch
    nm = lockstep.nelder_mead(ev, po["x"], lo, hi)
    assert np.all(nm["fun"] <= po["fun"]) and np.all((nm["x"] >= lo) & (nm["x"] <= hi))
    assert np.all(nm["nfev"] <= 15000) and np.all(nm["nit"] >= 1)
    ref = []
    for c in range(C):
        def f(x, c=c):
            x = np.asarray(x)
            return float(_rosen_like(x[None, :])[0]) if np.all((x >= lo[c]) & (x <= hi[c])) else np.inf
        p = minimize(f, list(x0[c]), method="Powell", options=dict(maxiter=12000, maxfev=12000, xtol=0.001, ftol=1))
        q = minimize(f, list(p.x), method="Nelder-Mead", options=dict(maxiter=15000, maxfev=15000, xatol=1e-4, fatol=1, adaptive=True))
        ref.append(q.fun)
    # same stopping rule (fatol = 1): both end within a few units of the minimum of the box
    assert np.median(nm["fun"]) <= np.median(ref) + 1.0 and np.max(nm["fun"] - np.array(ref)) < 5.0
    # the whole driver: DE generations batched over the chains, then the lock-step stages
    res, info = de_pw_nm(_PS(4), 6, seed=5, evaluate=_rosen_like)
    assert len(res) == 6 and all(s[2] <= s[1] <= s[0] for s in info["stages"])
    assert max(info["batches"]) == 6 * 30 * 4 and np.median(info["batches"]) >= 6


def test_opt_files_have_the_reference_layout(tmp_path):
    """The files must be readable the way the reference reads them (np.genfromtxt, header with
    '#Chi2 <names>', chi2 first: Examples/miscellany.ipynb cell 7)."""
    from nauyaca_b200.optimizers import write_opt_files

    class PS:
        system_name = "SysT"
        params_names = "mass1  period1  ecc1  mass2"
    rows = [[12.5, [0.1, 0.2, 0.3, 0.4], [1.0, 5.35, 0.01, 2.0]], [3.25, [0.5, 0.6, 0.7, 0.8], [3.0, 5.36, 0.02, 4.0]]]
    c, p = write_opt_files(PS, rows, path=str(tmp_path) + "/")
    assert c.endswith("SysT_cube.opt") and p.endswith("SysT_phys.opt")
    assert open(c).readline().split() == ["#Chi2", "mass1", "period1", "ecc1", "mass2"]
    cube = np.genfromtxt(c); phys = np.genfromtxt(p)
    assert cube.shape == (2, 5) and np.array_equal(cube[:, 0], [12.5, 3.25]) and np.array_equal(cube[1, 1:], [0.5, 0.6, 0.7, 0.8])
    assert np.array_equal(phys[0, 1:], [1.0, 5.35, 0.01, 2.0])
    best = list(sorted(cube, key=lambda x: x[0]))[0][1:]          # the notebook's idiom
    assert list(best) == [0.5, 0.6, 0.7, 0.8]
