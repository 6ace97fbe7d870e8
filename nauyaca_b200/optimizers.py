"""The optimizer seam of the reference (optimizers.py:46-203) on the batched GPU path.

``Optimizers.run`` starts ``nsols`` independent DE -> Powell -> Nelder-Mead chains, one
process each (optimizers.py:185-189); every chain calls ``calculate_chi2`` one vector at a
time, so the reference evaluates at most ``cores`` vectors concurrently. Two ways to run the
chains on the batched objective (``de_pw_nm(method=...)``):

``"lockstep"`` (default): the DE generation of every chain in one batch (30*ndim vectors per
chain), then Powell and Nelder-Mead of ALL chains advanced together (lockstep.py): each move
of the search is one ``ttvb_loglike`` batch over the live chains, with the reference's options
and stopping rules per chain.

``"scipy"``: the SAME scipy calls as optimizers.py:78-113, one per Python thread, their objective
calls meeting in a *rendezvous*: all pending vectors are evaluated in one batch, as soon as the
coordinator is free (default) or whenever every live chain is waiting for a value
(``greedy=False``). Same search path as the serial reference for the same seeds, but Powell and
Nelder-Mead contribute one vector per chain and round and the host work is serial Python.

``write_opt_files`` writes the two result tables in the reference's own text layout
(optimizers.py:122-136, 160-168; utils.writefile :1222) so that ``MCMC(opt_data=np.genfromtxt(
'*_cube.opt'))`` and ``init_walkers`` of the reference read them unchanged.
"""
import threading

import numpy as np

from . import utils as _u

__all__ = ["RendezvousBatcher", "de_pw_nm", "run_optimizers", "write_opt_files"]


class RendezvousBatcher:
    """Turns many blocking single-caller objectives into few large batches.

    worker threads call ``evaluate(X[m, ndim]) -> f[m]``; ``serve()`` (coordinator thread) waits
    until every live worker has a request pending, evaluates all of them with one
    ``batch_fn(X[M, ndim]) -> f[M]`` call and hands the slices back. Every worker sleeps on its
    own Event (no thundering herd on a shared condition variable)."""

    def __init__(self, batch_fn, nworkers, greedy=True):
        """greedy: a batch goes out as soon as the coordinator is free and anything is pending
        (the launch in flight hides the host work of the other chains; launches are latency-bound,
        so their size does not matter), instead of waiting until every live chain has asked."""
        self.batch_fn = batch_fn
        self.greedy = bool(greedy)
        self.live = int(nworkers)
        self.lock = threading.Lock()
        self.ready = threading.Event()     # "every live worker is waiting" or "a worker left"
        self.pending = []                  # [X, holder]
        self.batches = []                  # sizes of the batches served (diagnostics)
        self.error = None
        self._local = threading.local()

    def evaluate(self, X):
        X = np.atleast_2d(np.asarray(X, dtype=np.float64))
        ev = getattr(self._local, "ev", None)
        if ev is None:
            ev = self._local.ev = threading.Event()
        ev.clear()
        holder = {"f": None, "ev": ev}
        with self.lock:
            if self.error is not None:
                raise RuntimeError("batch evaluation failed") from self.error
            self.pending.append((X, holder))
            if self.greedy or len(self.pending) >= self.live:
                self.ready.set()
        ev.wait()
        if holder["f"] is None:
            raise RuntimeError("batch evaluation failed") from self.error
        return holder["f"]

    def worker_done(self):
        with self.lock:
            self.live -= 1
            self.ready.set()

    def serve(self):
        while True:
            self.ready.wait()
            with self.lock:
                self.ready.clear()
                if self.live <= 0 and not self.pending:
                    return
                if not self.pending or (not self.greedy and len(self.pending) < self.live):
                    continue
                todo, self.pending = self.pending, []
            try:
                sizes = [x.shape[0] for x, _ in todo]
                f = np.asarray(self.batch_fn(np.concatenate([x for x, _ in todo], axis=0)), dtype=np.float64)
                self.batches.append(int(sum(sizes)))
            except BaseException as e:      # wake the workers up, then re-raise here
                with self.lock:
                    self.error = e
                    rest, self.pending = self.pending, []
                for _, holder in todo + rest:
                    holder["ev"].set()
                raise
            o = 0
            for (x, holder), n in zip(todo, sizes):
                holder["f"] = f[o:o + n]
                o += n
                holder["ev"].set()


def _chain(PSystem_bounds, ndim, evaluate, seed, de_opts, powell_opts, nm_opts, out, idx, batcher):
    """One DE -> Powell -> Nelder-Mead chain with the reference's settings (optimizers.py:78-113)."""
    from scipy.optimize import differential_evolution, minimize
    try:
        box = {"lo": np.array([b[0] for b in PSystem_bounds]), "hi": np.array([b[1] for b in PSystem_bounds])}

        def chi2_block(X):
            # calculate_chi2 (utils.py:244-246): +inf outside the chain's current hypercube
            X = np.atleast_2d(X)
            inside = np.all((X >= box["lo"]) & (X <= box["hi"]), axis=1)
            f = np.full(X.shape[0], np.inf)
            if inside.any():
                f[inside] = evaluate(X[inside])
            return f

        def f_vec(x):                      # scipy vectorised DE: x is (ndim, S)
            x = np.asarray(x)
            if x.ndim == 1:
                return float(chi2_block(x[None, :])[0])
            return chi2_block(np.ascontiguousarray(x.T))

        def f_one(x):
            return float(chi2_block(np.asarray(x)[None, :])[0])

        de = differential_evolution(f_vec, list(zip(box["lo"], box["hi"])), seed=seed, vectorized=True,
                                    updating="deferred", **de_opts)
        x0, f0 = de.x, de.fun
        # optimizers.py:90: shrink the hypercube around the DE solution for the local stages
        box["lo"] = np.maximum(0.0, x0 - 0.3)
        box["hi"] = np.minimum(1.0, x0 + 0.3)
        x1, f1 = x0, f0
        if powell_opts is not None:
            po = minimize(f_one, list(x0), method="Powell", options=powell_opts)
            x1, f1 = po.x, po.fun
        nm = minimize(f_one, list(x1), method="Nelder-Mead", options=nm_opts)
        out[idx] = (float(f0), float(f1), float(nm.fun), np.asarray(nm.x, dtype=np.float64))
    except BaseException as e:
        out[idx] = e
    finally:
        batcher.worker_done()


def _de_only(PSystem_bounds, evaluate, seed, de_opts, out, idx, batcher):
    """The differential-evolution stage of one chain (optimizers.py:78-83)."""
    from scipy.optimize import differential_evolution
    try:
        lo = np.array([b[0] for b in PSystem_bounds]); hi = np.array([b[1] for b in PSystem_bounds])

        def f_vec(x):
            x = np.asarray(x)
            X = np.atleast_2d(x) if x.ndim == 1 else np.ascontiguousarray(x.T)
            inside = np.all((X >= lo) & (X <= hi), axis=1)
            f = np.full(X.shape[0], np.inf)
            if inside.any():
                f[inside] = evaluate(X[inside])
            return float(f[0]) if x.ndim == 1 else f
        de = differential_evolution(f_vec, list(zip(lo, hi)), seed=seed, vectorized=True, updating="deferred", **de_opts)
        out[idx] = (np.asarray(de.x, dtype=np.float64), float(de.fun))
    except BaseException as e:
        out[idx] = e
    finally:
        batcher.worker_done()


def de_pw_nm(PSystem, nsols, seed=None, evaluate=None, powell=True, de_opts=None, powell_opts=None, nm_opts=None,
             greedy=True, method="lockstep"):
    """``nsols`` DE -> Powell -> Nelder-Mead chains on the batched objective.

    Returns a list of ``[f, x_cube(list), x_phys(list)]`` like ``Optimizers.DE_PW_NM``
    (optimizers.py:46-139), in chain order, plus the batch-size log as second value.
    evaluate: X[B, ndim] -> chi2[B] on cube vectors (default: a private GPU engine of PSystem with the
    full unit hypercube; each chain applies its own, shrunken, bounds on the host).
    method: "lockstep" (default) - the DE generations of all chains in one batch each, then Powell
    and Nelder-Mead of all chains advanced together (nauyaca_b200/lockstep.py: every move of the
    search is ONE batch over the live chains); "scipy" - the reference's own scipy calls, one
    chain per thread, their objective calls meeting in rendezvous batches (same search path as
    the serial reference for the same seeds; an order of magnitude more launches)."""
    ndim = PSystem.ndim
    own_engine = None
    if evaluate is None:
        from .engine import BatchEngine
        own_engine = eng = BatchEngine(PSystem, device=_u.DEFAULT_DEVICE)     # private: its cube bounds are the unit
        eng.set_cube_bounds([[0.0, 1.0]] * ndim)                              # hypercube, whatever PSystem.hypercube is

        def evaluate(X):
            return eng.loglike(np.ascontiguousarray(X), mode=0)[0]
    de_o = dict(disp=False, popsize=30, maxiter=1, polish=False, mutation=(1.2, 1.7), recombination=0.6)
    de_o.update(de_opts or {})
    pw_o = dict(maxiter=12000, maxfev=12000, xtol=0.001, ftol=1, disp=False)
    pw_o.update(powell_opts or {})
    nm_o = dict(maxiter=15000, maxfev=15000, xatol=0.0001, fatol=1, disp=False, adaptive=True)
    nm_o.update(nm_opts or {})
    ss = np.random.SeedSequence(seed)
    seeds = [int(s.generate_state(1)[0]) for s in ss.spawn(nsols)]
    hyper = [list(h) for h in PSystem.hypercube]
    try:
        if method == "scipy":
            batcher = RendezvousBatcher(evaluate, nsols, greedy=greedy)
            out = [None] * nsols
            threads = [threading.Thread(target=_chain, daemon=True,
                                        args=(hyper, ndim, batcher.evaluate, seeds[i], de_o,
                                              pw_o if powell else None, nm_o, out, i, batcher)) for i in range(nsols)]
            for t in threads:
                t.start()
            batcher.serve()
            for t in threads:
                t.join()
            for r in out:
                if isinstance(r, BaseException):
                    raise r
            batches = batcher.batches
        elif method == "lockstep":
            from . import lockstep
            # DE: one thread per chain, every generation of all chains in one batch
            batcher = RendezvousBatcher(evaluate, nsols, greedy=False)
            de = [None] * nsols
            threads = [threading.Thread(target=_de_only, daemon=True, args=(hyper, batcher.evaluate, seeds[i], de_o, de, i, batcher))
                       for i in range(nsols)]
            for t in threads:
                t.start()
            batcher.serve()
            for t in threads:
                t.join()
            for r in de:
                if isinstance(r, BaseException):
                    raise r
            x0 = np.stack([r[0] for r in de]); f0 = np.array([r[1] for r in de])
            # optimizers.py:90: shrink the hypercube around the DE solution for the local stages
            lo, hi = np.maximum(0.0, x0 - 0.3), np.minimum(1.0, x0 + 0.3)
            batches = list(batcher.batches)
            x1, f1 = x0, f0
            if powell:
                po = lockstep.powell(evaluate, x0, lo, hi, f0=f0, xtol=pw_o["xtol"], ftol=pw_o["ftol"],
                                     maxiter=pw_o["maxiter"], maxfev=pw_o["maxfev"])
                x1, f1 = po["x"], po["fun"]
                batches += po["batches"]
            nm = lockstep.nelder_mead(evaluate, x1, lo, hi, xatol=nm_o["xatol"], fatol=nm_o["fatol"], maxiter=nm_o["maxiter"],
                                      maxfev=nm_o["maxfev"], adaptive=nm_o["adaptive"])
            batches += nm["batches"]
            out = [(float(f0[i]), float(f1[i]), float(nm["fun"][i]), nm["x"][i].copy()) for i in range(nsols)]
        else:
            raise ValueError("method must be 'lockstep' or 'scipy'")
        X = np.stack([r[3] for r in out])
        phys = None
        try:
            flat, _ = _u.engine_for(PSystem).cube_to_physical(X)
            phys = [list(_u._remove_constants(PSystem, row)) for row in flat]
        except Exception:
            phys = [None] * nsols
    finally:
        if own_engine is not None:
            own_engine.close()
    results = [[out[i][2], X[i].tolist(), phys[i]] for i in range(nsols)]
    return results, {"batches": batches, "stages": [(r[0], r[1], r[2]) for r in out]}


def run_optimizers(PSystem, nsols=1, seed=None, **kw):
    """``Optimizers.run`` seam (optimizers.py:142-203) without the files: dict with 'chi2',
    'cube', 'physical' ordered by chi2 like ``Optimizers.results``."""
    res, info = de_pw_nm(PSystem, nsols, seed=seed, **kw)
    order = np.argsort([r[0] for r in res])
    return {"chi2": np.array([res[i][0] for i in order]), "cube": np.array([res[i][1] for i in order]),
            "physical": np.array([res[i][2] for i in order]), "info": info}


def _writefile(path, mode, text, align):
    """utils.py:1222-1232 writefile(): text.split() formatted with `align`."""
    with open(path, mode) as fh:
        fh.write(align % tuple(text.split()))


def write_opt_files(PSystem, results, path="./", suffix=""):
    """Write ``<system_name>_cube<suffix>.opt`` and ``<system_name>_phys<suffix>.opt`` in the layout
    of the reference: a '#Chi2 <params_names>' header (optimizers.py:160-166), then one line per
    solution, chi2 followed by the values (optimizers.py:122-136). `results` is the list returned
    by de_pw_nm or the dict returned by run_optimizers. Returns the two file names."""
    if isinstance(results, dict):
        rows = [[c, list(x), list(p)] for c, x, p in zip(results["chi2"], results["cube"], results["physical"])]
    else:
        rows = results
    base_cube = f"{path}{PSystem.system_name}_cube{suffix}.opt"
    base_phys = f"{path}{PSystem.system_name}_phys{suffix}.opt"
    header = "#Chi2 " + " ".join(PSystem.params_names.split())
    for name in (base_cube, base_phys):
        _writefile(name, "w", header, "%-15s" + " %-11s" * (len(header.split()) - 1) + "\n")
    for f2, x_cube, x_phys in rows:
        info = "{} \t {} \n".format(str(f2), "  ".join(str(v) for v in x_cube))
        _writefile(base_cube, "a", info, "%-30s" + " %-11s" * (len(info.split()) - 1) + "\n")
        info = " " + str(f2) + " " + "  ".join(str(v) for v in x_phys)
        _writefile(base_phys, "a", info, "%-30s" + " %-11s" * (len(info.split()) - 1) + "\n")
    return base_cube, base_phys
