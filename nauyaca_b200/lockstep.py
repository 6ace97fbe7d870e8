"""Local search of MANY optimizer chains in lock-step.

The reference refines every differential-evolution solution with scipy's Powell and then scipy's
Nelder-Mead (optimizers.py:95-111), one chain per process, one objective call at a time. On the
GPU a launch costs the same ~1.5 ms for 1 vector as for 30 000, so the chains are advanced
TOGETHER here: every move of the search asks for the candidates of all live chains in one batch.

``powell``       Powell's conjugate-direction method; its line minimisations evaluate a comb of
                 bracketing points per chain and round (refined until the spacing is below xtol)
                 instead of a serial Brent search.
``nelder_mead``  the adaptive simplex of Gao & Han (scipy's ``adaptive=True``); reflection, expansion
                 and both contractions of a chain are evaluated speculatively in ONE batch, a shrink
                 in a second one.
Both keep the reference's options and stopping rules (xtol / ftol, xatol / fatol, maxiter, maxfev)
per chain; chains that have stopped leave the batches. Infeasible points (outside a chain's box)
cost +inf, as ``calculate_chi2`` returns (utils.py:244-246).

The search paths are not those of the serial scipy calls (a comb is not Brent, and the objective
is deterministic, so the RESULTS are reproducible but not identical): DESIGN.md records the
comparison of the final chi2 distributions. ``optimizers.de_pw_nm(method="scipy")`` keeps the
one-thread-per-chain scipy path.
"""
import numpy as np

__all__ = ["powell", "nelder_mead", "BoxedObjective"]


class BoxedObjective:
    """f(X[C, m, n]) -> F[C, m]: one batch for all chains; +inf outside each chain's box [lo, hi]."""

    def __init__(self, evaluate, lo, hi):
        self.evaluate, self.lo, self.hi = evaluate, np.asarray(lo, float), np.asarray(hi, float)
        self.nfev = np.zeros(self.lo.shape[0], dtype=np.int64)
        self.batches = []

    def __call__(self, X, rows):
        """X[R, m, n] for the chains `rows` (indices into the box arrays) -> F[R, m]."""
        X = np.asarray(X, float)
        R, m, n = X.shape
        inside = np.all((X >= self.lo[rows, None, :]) & (X <= self.hi[rows, None, :]), axis=2)
        F = np.full((R, m), np.inf)
        if inside.any():
            F[inside] = np.asarray(self.evaluate(np.ascontiguousarray(X[inside])), float)
            self.batches.append(int(inside.sum()))
        F[np.isnan(F)] = np.inf
        self.nfev[rows] += m
        return F


# ------------------------------------------------------------------------------------ Powell
def _line_min(obj, rows, x, fx, d, xtol, comb=8):
    """Minimise f(x + a d) over a for the chains `rows`: combs of `comb` points inside the feasible
    interval, recentred on the best point and narrowed until the spacing is below xtol (in units of
    the largest component of d). Returns (x_new, f_new); a chain never moves uphill."""
    R, n = x.shape
    with np.errstate(divide="ignore", invalid="ignore"):
        t1 = (obj.lo[rows] - x) / d
        t2 = (obj.hi[rows] - x) / d
    tlo = np.where(d != 0, np.minimum(t1, t2), -np.inf).max(axis=1)
    thi = np.where(d != 0, np.maximum(t1, t2), np.inf).min(axis=1)
    dead = ~np.isfinite(tlo) | ~np.isfinite(thi) | (thi <= tlo)
    tlo = np.where(dead, 0.0, np.minimum(tlo, 0.0)); thi = np.where(dead, 0.0, np.maximum(thi, 0.0))
    scale = np.maximum(np.abs(d).max(axis=1), 1e-300)
    a_best, f_best = np.zeros(R), fx.copy()
    lo_a, hi_a = tlo.copy(), thi.copy()
    live = ~dead
    while live.any():
        idx = np.nonzero(live)[0]
        frac = (np.arange(comb) + 1.0) / (comb + 1.0)
        A = lo_a[idx, None] + (hi_a[idx] - lo_a[idx])[:, None] * frac[None, :]
        F = obj(x[idx, None, :] + A[:, :, None] * d[idx, None, :], rows[idx])
        k = np.argmin(F, axis=1)
        fk = F[np.arange(idx.size), k]
        better = fk < f_best[idx]
        a_best[idx] = np.where(better, A[np.arange(idx.size), k], a_best[idx])
        f_best[idx] = np.where(better, fk, f_best[idx])
        h = (hi_a[idx] - lo_a[idx]) / (comb + 1.0)
        lo_a[idx] = np.maximum(tlo[idx], a_best[idx] - h)
        hi_a[idx] = np.minimum(thi[idx], a_best[idx] + h)
        live[idx] = (h * scale[idx]) > xtol
    return x + a_best[:, None] * d, f_best


def powell(evaluate, x0, lo, hi, f0=None, xtol=1e-3, ftol=1.0, maxiter=12000, maxfev=12000, comb=8):
    """Powell's method for C chains at once. x0[C, n]; lo, hi[C, n] the chains' boxes.
    Returns dict(x[C, n], fun[C], nfev[C], nit[C], batches)."""
    x = np.array(x0, float)
    C, n = x.shape
    obj = BoxedObjective(evaluate, lo, hi)
    allrows = np.arange(C)
    fval = obj(x[:, None, :], allrows)[:, 0] if f0 is None else np.array(f0, float)
    direc = np.tile(np.eye(n), (C, 1, 1))
    nit = np.zeros(C, dtype=np.int64)
    active = np.isfinite(fval)
    while active.any():
        rows = np.nonzero(active)[0]
        x1, fx = x[rows].copy(), fval[rows].copy()
        delta, bigind = np.zeros(rows.size), np.zeros(rows.size, dtype=np.int64)
        for i in range(n):
            f_before = fval[rows].copy()
            xn, fn = _line_min(obj, rows, x[rows], fval[rows], direc[rows, i, :], xtol, comb)
            x[rows], fval[rows] = xn, fn
            gain = f_before - fn
            up = gain > delta
            delta = np.where(up, gain, delta); bigind = np.where(up, i, bigind)
        nit[rows] += 1
        fv = fval[rows]
        stop = 2.0 * (fx - fv) <= ftol * (np.abs(fx) + np.abs(fv)) + 1e-20          # scipy's test (ftol is relative)
        stop |= (obj.nfev[rows] >= maxfev) | (nit[rows] >= maxiter)
        go = ~stop
        if go.any():
            r = rows[go]
            d1 = x[r] - x1[go]
            x2 = 2.0 * x[r] - x1[go]
            fx2 = obj(x2[:, None, :], r)[:, 0]
            fxg, fvg, dl = fx[go], fv[go], delta[go]
            with np.errstate(invalid="ignore"):
                t = 2.0 * (fxg + fx2 - 2.0 * fvg) * (fxg - fvg - dl) ** 2 - dl * (fxg - fx2) ** 2
                ext = (fxg > fx2) & (t < 0.0)
            if ext.any():
                re = r[ext]
                xn, fn = _line_min(obj, re, x[re], fval[re], d1[ext], xtol, comb)
                moved = np.any(xn != x[re], axis=1)
                x[re], fval[re] = xn, fn
                bi = bigind[go][ext]
                for c, b, m, dd in zip(re, bi, moved, d1[ext]):
                    if m:
                        direc[c, b, :] = direc[c, -1, :]
                        direc[c, -1, :] = dd
        active[rows[stop]] = False
    return {"x": x, "fun": fval, "nfev": obj.nfev.copy(), "nit": nit, "batches": obj.batches}


# ------------------------------------------------------------------------------------ Nelder-Mead
def nelder_mead(evaluate, x0, lo, hi, xatol=1e-4, fatol=1.0, maxiter=15000, maxfev=15000, adaptive=True):
    """Adaptive Nelder-Mead for C chains at once (scipy's algorithm and start simplex).
    Returns dict(x[C, n], fun[C], nfev[C], nit[C], batches)."""
    x0 = np.array(x0, float)
    C, n = x0.shape
    obj = BoxedObjective(evaluate, lo, hi)
    if adaptive:
        rho, chi, psi, sigma = 1.0, 1.0 + 2.0 / n, 0.75 - 1.0 / (2.0 * n), 1.0 - 1.0 / n
    else:
        rho, chi, psi, sigma = 1.0, 2.0, 0.5, 0.5
    sim = np.repeat(x0[:, None, :], n + 1, axis=1)
    for k in range(n):
        col = sim[:, k + 1, k]
        sim[:, k + 1, k] = np.where(col != 0.0, 1.05 * col, 0.00025)
    allrows = np.arange(C)
    fsim = obj(sim, allrows)
    nit = np.zeros(C, dtype=np.int64)
    order = np.argsort(fsim, axis=1, kind="stable")
    fsim = np.take_along_axis(fsim, order, axis=1)
    sim = np.take_along_axis(sim, order[:, :, None], axis=1)
    active = np.ones(C, dtype=bool)
    while True:
        with np.errstate(invalid="ignore"):
            conv = (np.abs(sim[:, 1:, :] - sim[:, :1, :]).max(axis=(1, 2)) <= xatol) & \
                   (np.abs(fsim[:, 1:] - fsim[:, :1]).max(axis=1) <= fatol)
        active &= ~conv & (obj.nfev < maxfev) & (nit < maxiter)
        if not active.any():
            break
        r = np.nonzero(active)[0]
        S, Fs = sim[r], fsim[r]
        xbar = S[:, :-1, :].mean(axis=1)
        worst = S[:, -1, :]
        cand = np.stack([(1 + rho) * xbar - rho * worst,                       # reflection
                         (1 + rho * chi) * xbar - rho * chi * worst,           # expansion
                         (1 + psi * rho) * xbar - psi * rho * worst,           # outside contraction
                         (1 - psi) * xbar + psi * worst], axis=1)              # inside contraction
        Fc = obj(cand, r)
        fr, fe, fc, fcc = Fc[:, 0], Fc[:, 1], Fc[:, 2], Fc[:, 3]
        f0, fn1, fn = Fs[:, 0], Fs[:, -2], Fs[:, -1]
        # scipy counts only the evaluations the serial algorithm would have made
        used = np.where(fr < f0, 2, np.where(fr < fn1, 1, 2))
        obj.nfev[r] += used - 4
        take = np.full(r.size, -1)                                             # index into cand, -1 = shrink
        take = np.where(fr < f0, np.where(fe < fr, 1, 0), take)
        mid = ~(fr < f0) & (fr < fn1)
        take = np.where(mid, 0, take)
        hi_ = ~(fr < f0) & ~(fr < fn1)
        take = np.where(hi_ & (fr < fn) & (fc <= fr), 2, take)
        take = np.where(hi_ & ~(fr < fn) & (fcc < fn), 3, take)
        acc = take >= 0
        if acc.any():
            ia = np.nonzero(acc)[0]
            S[ia, -1, :] = cand[ia, take[ia], :]
            Fs[ia, -1] = Fc[ia, take[ia]]
        if (~acc).any():
            ish = np.nonzero(~acc)[0]
            S[ish, 1:, :] = S[ish, :1, :] + sigma * (S[ish, 1:, :] - S[ish, :1, :])
            Fs[ish, 1:] = obj(S[ish, 1:, :], r[ish])
        order = np.argsort(Fs, axis=1, kind="stable")
        fsim[r] = np.take_along_axis(Fs, order, axis=1)
        sim[r] = np.take_along_axis(S, order[:, :, None], axis=1)
        nit[r] += 1
    return {"x": sim[:, 0, :].copy(), "fun": fsim[:, 0].copy(), "nfev": obj.nfev.copy(), "nit": nit,
            "batches": obj.batches}
