"""CPU: the kernel's arithmetic header (nauyaca_b200/csrc/ttvb_math.cuh) compiled for the host
by the TEST-ONLY harness must equal the independent oracle bit for bit."""
import numpy as np
import pytest


@pytest.mark.parametrize("name,n", [("doc2", 60), ("sys3", 60), ("sys3f", 15)])
def test_harness_equals_oracle_bitwise(oracle, harness, systems, name, n):
    s = systems[name]; osys = oracle.OracleSystem(s)
    rng = np.random.default_rng(11)
    nev = 0
    for _ in range(n):
        flat = osys.cube_to_physical(rng.random(s["ndim"]))
        a = oracle.ephemeris(flat, s["mstar"], s["t0"], s["dt"], s["ftime"])
        b = harness(flat, s["mstar"], s["t0"], s["dt"], s["ftime"])
        assert a[0] == b[0]
        for u, v in zip(a[1:], b[1:]):
            assert np.array_equal(u, v)
        nev += len(a[3])
    assert nev > 10 * n


def test_harness_golden(harness, systems, known):
    k = known["G1"]; s = systems["sys3"]
    st, pl, ep, tm, rs, vs = harness(k["flat_params"], s["mstar"], s["t0"], s["dt"], s["ftime"])
    assert st == 0 and len(tm) == 105
    for pid, idx in s["planets_IDs"].items():
        m = pl == idx
        assert np.max(np.abs(tm[m] - np.array(k["ephemeris"][pid]["times"]))) <= 1e-10
