"""CPU: known answers that need no reference-held golden - two-body transit times in closed form, the
massless limit, convergence under step halving. They pin the conventions of the prologue (elements ->
state, observer on +z), the Kepler drift and the transit locator independently
of the notebook values, for the oracle and for the host build of the kernel's arithmetic alike."""
import numpy as np
import pytest

G_TTV = 0.000295994511
MEARTH = 3.0034893488507934e-06


def _analytic_transits(P, e, w_deg, M0_deg, t0, tend):
    """Mid-transit times of a two-body orbit seen edge-on (i = 90 deg, observer on +z): the projected
    separation r |cos(w + f)| vanishes with z > 0 at w + f = 90 deg."""
    f = np.deg2rad(90.0 - w_deg)
    E = 2.0 * np.arctan2(np.sqrt(1.0 - e) * np.sin(f / 2.0), np.sqrt(1.0 + e) * np.cos(f / 2.0))
    M = E - e * np.sin(E)
    n = 2.0 * np.pi / P
    t = t0 + ((M - np.deg2rad(M0_deg)) % (2.0 * np.pi)) / n
    out = []
    while t < tend:
        out.append(t)
        t += P
    return np.array(out)


CASES = [  # mass [Mearth], P, e, w, M0, Omega
    (1.0, 3.7, 0.0, 90.0, 10.0, 0.0),
    (30.0, 11.3, 0.3, 25.0, 200.0, 40.0),
    (300.0, 7.77, 0.6, 310.0, 5.0, 190.0),
]


@pytest.mark.parametrize("which", ["oracle", "harness"])
@pytest.mark.parametrize("case", CASES)
def test_two_body_transit_times_in_closed_form(oracle, harness, which, case):
    mass, P, e, w, M0, Om = case
    mstar, t0, dt, tend = 0.8, 2.5, 0.05, 120.0
    flat = np.array([mass, P, e, 90.0, w, M0, Om])
    if which == "oracle":
        st, pl, ep, tm, rs, vs = oracle.ephemeris(flat, mstar, t0, dt, tend)
    else:
        st, pl, ep, tm, rs, vs = harness(flat, mstar, t0, dt, tend)
    assert st == 0
    want = _analytic_transits(P, e, w, M0, t0, tend)
    # an event needs the step before and the step after it inside the integration
    want = want[(want > t0 + dt) & (want < tend - dt)]
    got = np.asarray(tm)
    got = got[(got > t0 + dt) & (got < tend - dt)]
    assert len(got) == len(want) and len(want) >= 8
    assert np.max(np.abs(got - want)) < 2e-10, np.max(np.abs(got - want))
    assert np.all(np.diff(np.asarray(ep)) == 1)
    # edge-on: the sky-projected separation at mid-transit is zero (to round-off of the orbit size)
    a = (G_TTV * (mstar + mass * MEARTH) * P * P / (4 * np.pi ** 2)) ** (1.0 / 3.0)
    assert np.max(np.abs(rs)) < 1e-7 * a            # sqrt of a round-off-sized x^2 + y^2


def test_massless_limit_is_two_independent_kepler_orbits(oracle):
    """Two planets of 1e-7 Earth masses: each keeps its own period to 1e-10 d over 150 d."""
    mstar, t0, dt, tend = 1.1, 0.0, 0.04, 150.0
    flat = np.array([1e-7, 5.3, 0.1, 90.0, 40.0, 100.0, 0.0, 1e-7, 8.9, 0.2, 90.0, 250.0, 30.0, 0.0])
    st, pl, ep, tm, rs, vs = oracle.ephemeris(flat, mstar, t0, dt, tend)
    assert st == 0
    for j, (P, e, w, M0) in enumerate([(5.3, 0.1, 40.0, 100.0), (8.9, 0.2, 250.0, 30.0)]):
        t = np.asarray(tm)[np.asarray(pl) == j]
        want = _analytic_transits(P, e, w, M0, t0, tend)
        want = want[(want > t0 + dt) & (want < tend - dt)]
        t = t[(t > t0 + dt) & (t < tend - dt)]
        assert len(t) == len(want)
        assert np.max(np.abs(t - want)) < 5e-9      # masses of 3e-13 Msun still move the star a little


def test_step_halving_converges_with_the_corrector(oracle, systems):
    """Interacting three-planet system (G1's): the transit times at dt, dt/2, dt/4 converge, and each
    halving cuts the difference by far more than the factor four of the bare second-order map: with the
    third-order symplectic corrector applied at the start the leading error term, of first
    order in the planet masses, falls as dt^4 (measured: a factor ~29 per halving; a missing or wrong
    corrector would give ~4)."""
    spec = systems["sys3"]
    from oracle import oracle_py
    osys = oracle_py.OracleSystem(spec)
    x = np.full(spec["ndim"], 0.5)
    flat = osys.cube_to_physical(x)
    runs = []
    for k in (1, 2, 4):
        st, pl, ep, tm, rs, vs = oracle.ephemeris(flat, spec["mstar"], spec["t0"], 0.2 / k, 200.0)
        assert st == 0
        runs.append({(int(p), int(e)): t for p, e, t in zip(pl, ep, tm)})
    keys = set(runs[0]) & set(runs[1]) & set(runs[2])
    assert len(keys) >= 30
    d1 = np.array([abs(runs[0][k] - runs[1][k]) for k in sorted(keys)])
    d2 = np.array([abs(runs[1][k] - runs[2][k]) for k in sorted(keys)])
    assert d1.max() < 1e-4 and d2.max() < d1.max()
    ratio = np.median(d1[d2 > 1e-13] / d2[d2 > 1e-13])
    assert 10.0 < ratio < 40.0, ratio
