"""`ttvfast`-shaped shim backed by the CPU oracle (TEST INFRASTRUCTURE).

Lets the UNMODIFIED reference nauyaca/utils.py run end to end in the development container,
where ttvfast==0.3.0 is not installable: `ttvfast.ttvfast(...)` has the call signature and
return layout utils.py:72-79 relies on (5 lists of 5000 padded with -2)."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from . import models  # noqa: E402,F401

_MEARTH = 3.0034893488507934e-06
N_EVENTS = 5000


def ttvfast(planets, stellar_mass, time, dt, total, rv_times=None, input_flag=1):
    from oracle import oracle_py
    if input_flag != 1 or rv_times is not None:
        raise NotImplementedError("shim covers the call nauyaca makes: input_flag=1, rv_times=None")
    flat = []
    for p in planets:
        # utils.py:62 multiplied by Mearth_to_Msun; undo with the same constant (exact for the
        # golden vectors is not required here: the oracle multiplies again, see tests)
        flat += [p.mass / _MEARTH, p.period, p.eccentricity, p.inclination, p.argument, p.mean_anomaly, p.longnode]
    # TTVO_NATIVE=1 (bench.py's reference-plumbing timing leg): the oracle build with the host's own
    # divide / sqrt instead of the emulated GPU seeds
    st, pl, ep, tm, rs, vs = oracle_py.ephemeris(flat, stellar_mass, time, dt, total, cap=N_EVENTS,
                                                 native=os.environ.get("TTVO_NATIVE") == "1")
    pad = [-2] * (N_EVENTS - len(tm))
    pos = [pl.tolist() + pad, ep.tolist() + pad, tm.tolist() + pad, rs.tolist() + pad, vs.tolist() + pad]
    return {"positions": pos, "rv": None}
