"""Host-side mirror of the reference's interface for the hot path (nauyaca/utils.py).

Same names, argument meaning and error behaviour as the reference functions, backed by the
GPU engine (BatchEngine) instead of ttvfast + Python loops. The single-vector functions
exist so that code written against nauyaca.utils keeps working; the ``*_batch`` functions
are the ones that make the GPU worthwhile. There is no CPU fallback.

reference function                     here
-------------------------------------  -----------------------------------------------
run_TTVFast            utils.py:16     run_TTVFast            (engine seam, table mode)
_ephemeris             utils.py:84     _ephemeris
calculate_ephemeris    utils.py:126    calculate_ephemeris
_chi2                  utils.py:161    _chi2
calculate_chi2         utils.py:219    calculate_chi2, calculate_chi2_batch
calculate_chi2_physical utils.py:260   calculate_chi2_physical
log_likelihood_func    utils.py:333    log_likelihood_func, log_likelihood_batch
cube_to_physical       utils.py:403    cube_to_physical
_insert_constants      utils.py:440    _insert_constants
_remove_constants      utils.py:450    _remove_constants
intervals              utils.py:1193   intervals
init_walkers           utils.py:463    init_walkers (walkers.py)
"""
import sys

import numpy as np

from ._lib import MODE_CUBE, MODE_PHYSICAL, MODE_PHYSICAL_FULL, EVENT_CAP
from .engine import BatchEngine, system_spec
from .walkers import init_walkers  # noqa: F401  (reference keeps it in utils)

__all__ = ['run_TTVFast', 'calculate_ephemeris', 'log_likelihood_func', 'cube_to_physical',
           '_ephemeris', '_remove_constants', '_insert_constants', 'calculate_chi2',
           'calculate_chi2_physical', 'intervals', 'calculate_chi2_batch', 'log_likelihood_batch',
           'engine_for', 'clear_engine_cache', 'init_walkers']

Mearth_to_Msun = 3.0034893488507934e-06     # reference constants.py:65

_ENGINES = {}
_RAW_ENGINES = {}
DEFAULT_DEVICE = 0


def _fingerprint(PS):
    tt = tuple(tuple(v.values()) for v in PS.transit_times.values())
    ss = tuple(tuple(v.values()) for v in PS.sigma_obs.values())
    ee = tuple(tuple(v.keys()) for v in PS.transit_times.values())
    return (PS.mstar, PS.rstarAU, PS.t0, PS.ftime, PS.dt, PS.npla, PS.ndim, tuple(PS.bi), tuple(PS.bf),
            tuple(map(tuple, PS.bounds)), tuple(PS.constant_params.items()), tt, ss, ee,
            tuple(PS.planets_IDs.items()))


def engine_for(PSystem, device=None):
    """The BatchEngine of a PlanetarySystem (cached; rebuilt when an attribute the path
    consumes changes; PSystem.hypercube is followed without a rebuild, optimizers.py:90)."""
    device = DEFAULT_DEVICE if device is None else device
    if isinstance(PSystem, BatchEngine):
        return PSystem
    fp = _fingerprint(PSystem)
    key = (id(PSystem), device)
    hit = _ENGINES.get(key)
    if hit is None or hit[0] != fp:
        if hit is not None:
            hit[1].close()
        hit = [fp, BatchEngine(PSystem, device=device), None]
        _ENGINES[key] = hit
    hc = tuple(map(tuple, PSystem.hypercube))
    if hit[2] != hc:
        hit[1].set_cube_bounds(PSystem.hypercube)
        hit[2] = hc
    return hit[1]


def clear_engine_cache():
    for v in list(_ENGINES.values()):
        v[1].close()
    for v in list(_RAW_ENGINES.values()):
        v.close()
    _ENGINES.clear()
    _RAW_ENGINES.clear()


def _raw_engine(npla, mstar, t0, ftime, dt, device=None):
    """Engine without observations or bounds, for run_TTVFast(flat, mstar, ...)."""
    device = DEFAULT_DEVICE if device is None else device
    key = (npla, float(mstar), float(t0), float(ftime), float(dt), device)
    eng = _RAW_ENGINES.get(key)
    if eng is None:
        spec = {"mstar": float(mstar), "rstarAU": 0.0, "t0": float(t0), "ftime": float(ftime), "dt": float(dt),
                "npla": npla, "ndim": 7 * npla, "bounds": [[-np.inf, np.inf]] * (7 * npla),
                "bi": [0.0] * (7 * npla), "bf": [1.0] * (7 * npla), "hypercube": [[0.0, 1.0]] * (7 * npla),
                "constant_params": {}, "observed": {}, "second_term_sum": 0.0}
        eng = BatchEngine(spec, device=device)
        _RAW_ENGINES[key] = eng
    return eng


# ------------------------------------------------------------------------------------
def run_TTVFast(flat_params, mstar, init_time=0., final_time=None, dt=None):
    """utils.py:16-81. Returns the (5, n_events) array [planet, epoch, time, rsky, vsky]."""
    if (len(flat_params) % 7) == 0:
        pass
    else:
        sys.exit('flat_params do not contains the necessary elements. ' +
                 'Seven parameters per planet must be given')
    npla = len(flat_params) // 7
    eng = _raw_engine(npla, mstar, init_time, final_time, dt)
    n, pl, ep, tm, rs, vs, st = eng.ephemeris(np.asarray(flat_params, dtype=np.float64)[None, :],
                                              mode=MODE_PHYSICAL_FULL, cap=EVENT_CAP)
    k = int(n[0])
    return np.array([pl[0, :k].astype(float), ep[0, :k].astype(float), tm[0, :k], rs[0, :k], vs[0, :k]])


def _ephemeris(PSystem, SP):
    """utils.py:84-123: per planet {epoch: time} of the events with rsky <= rstarAU (one pass over the
    event table; a malformed table prints the reference's warning and yields what was built)."""
    ephemeris = {}
    try:
        planet, epoch, time, rsky = (np.asarray(SP[k]) for k in range(4))
        transits = rsky <= PSystem.rstarAU
        for planet_id, number in PSystem.planets_IDs.items():
            sel = transits & (planet == number)
            ephemeris[planet_id] = {int(e): t for e, t in zip(epoch[sel], time[sel])}
    except Exception:
        print('Warning: Invalid proposal')
    return ephemeris


def calculate_ephemeris(PSystem, flat_params):
    """utils.py:126-158."""
    SP = run_TTVFast(flat_params, PSystem.mstar, init_time=PSystem.t0, final_time=PSystem.ftime, dt=PSystem.dt)
    return _ephemeris(PSystem, SP)


def _chi2(observed, sigma, simulated, individual=False):
    """utils.py:161-216 on host dictionaries (the GPU computes the same sum in-kernel): per planet, in
    the order of its observed epochs, ((t_obs - t_sim) / sigma)^2, and 1e20 for an observed epoch the
    simulation does not have; the planets' sums are added in dictionary order."""
    per_planet = {}
    for pid, table in observed.items():
        model, err = simulated[pid], sigma[pid]
        acc = 0.0
        for ep, t_obs in table.items():
            t_sim = model.get(ep)
            acc += 1e+20 if t_sim is None else ((t_obs - t_sim) / err[ep]) ** 2
        per_planet[pid] = acc
    if individual:
        return per_planet
    total = 0.0
    for v in per_planet.values():
        total += v
    return total


def intervals(frontiers, flat_params):
    """utils.py:1193-1219: one boolean per parameter, cut short after the first False (callers only
    ask whether False is in the list)."""
    TF = []
    for (lo, hi), value in zip(frontiers, flat_params):
        TF.append(bool(lo <= value <= hi))
        if not TF[-1]:
            break
    if len(frontiers) < len(flat_params) and (not TF or TF[-1]):
        frontiers[len(frontiers)]           # the reference indexes past the bounds here: same IndexError
    return TF


def _insert_constants(PSystem, x):
    """utils.py:440-447: the constants go back to their flat indices, in ascending order."""
    out = list(x)
    for index in sorted(PSystem.constant_params):
        out.insert(index, PSystem.constant_params[index])
    return out


def _remove_constants(PSystem, x):
    """utils.py:450-460: drop the slots of the constant parameters."""
    skip = set(PSystem.constant_params.keys())
    return np.array([v for k, v in enumerate(x) if k not in skip])


def cube_to_physical(PSystem, x):
    """utils.py:403-438 (one vector; evaluated by the GPU kernel ttvb_cube_to_physical)."""
    flat, _inside = engine_for(PSystem).cube_to_physical(np.asarray(x, dtype=np.float64)[None, :])
    return flat[0]


# ------------------------------------------------------------------------------------ batch
def calculate_chi2_batch(X_cube, PSystem):
    """calculate_chi2 (utils.py:219-257) for every row of X_cube[B, ndim]: chi2[B], +inf
    outside PSystem.hypercube."""
    chi2, _logl, _st = engine_for(PSystem).loglike(X_cube, mode=MODE_CUBE)
    return chi2


def log_likelihood_batch(X, PSystem, cube=True, insert_constants=False):
    """log_likelihood_func (utils.py:333-399) for every row of X: logl[B]."""
    eng = engine_for(PSystem)
    if cube:
        return eng.loglike(X, mode=MODE_CUBE)[1]
    X = np.atleast_2d(np.asarray(X, dtype=np.float64))
    if insert_constants:
        return eng.loglike(X, mode=MODE_PHYSICAL)[1]
    # utils.py:379: intervals(PSystem.bounds, flat_params) on the FULL vector compares slot i
    # of the 7*npla vector with bounds[i] of the ndim free parameters (index-misaligned when
    # constants exist) and raises IndexError past len(bounds). Reproduced as is.
    out = np.empty(X.shape[0])
    keep = []
    for b, row in enumerate(X):
        if False in intervals(PSystem.bounds, row):
            out[b] = -np.inf
        else:
            keep.append(b)
    if keep:
        out[keep] = eng.loglike(X[keep], mode=MODE_PHYSICAL_FULL)[1]
    return out


# ------------------------------------------------------------------------------------ single
def calculate_chi2(flat_params, PSystem):
    """utils.py:219-257."""
    return float(calculate_chi2_batch(np.asarray(flat_params, dtype=np.float64)[None, :], PSystem)[0])


def calculate_chi2_physical(flat_params, PSystem, individual=False, insert_constants=False, get_ephemeris=False):
    """utils.py:260-330."""
    x = list(flat_params)
    if insert_constants:
        x = _insert_constants(PSystem, x)
    if len(x) == 7 * PSystem.npla:
        pass
    else:
        sys.exit("Length of -flat_params- mismatch number of dimensions per planet")
    eng = engine_for(PSystem)
    chi2, _logl, _st, per = eng.loglike(np.asarray(x, dtype=np.float64)[None, :], mode=MODE_PHYSICAL_FULL,
                                        per_planet=True)
    if individual:
        result = {pid: float(per[0, PSystem.planets_IDs[pid]]) for pid in PSystem.transit_times}
    else:
        result = float(chi2[0])
    if get_ephemeris:
        return (result, calculate_ephemeris(PSystem, x))
    return result


def log_likelihood_func(flat_params, PSystem, cube=True, insert_constants=False):
    """utils.py:333-399."""
    return float(log_likelihood_batch(np.asarray(flat_params, dtype=np.float64)[None, :], PSystem, cube=cube,
                                      insert_constants=insert_constants)[0])
