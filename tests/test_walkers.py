"""init_walkers mirror (nauyaca_b200/walkers.py) against golden populations produced by the
reference's own init_walkers (tests/golden/make_init_walkers.py): same numpy seed -> same bits."""
import json
import os
import types

import numpy as np
import pytest

from nauyaca_b200.walkers import init_walkers

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def golden():
    with open(os.path.join(HERE, "golden", "init_walkers.json")) as fh:
        return json.load(fh)


def _inputs(g):
    chi2 = np.array([float.fromhex(v) for v in g["chi2"]])
    cube = np.array([[float.fromhex(v) for v in row] for row in g["cube_hex"]])
    PS = types.SimpleNamespace(ndim=g["ndim"], hypercube=g["hypercube"], bounds=[[0.0, 1.0]] * g["ndim"])
    return PS, chi2, cube


def test_populations_equal_the_reference_bit_for_bit(golden, capsys):
    PS, chi2, cube = _inputs(golden)
    for c in golden["cases"]:
        opt = None
        if c["distribution"] != "Uniform":
            opt = {"chi2": chi2.copy(), "cube": cube.copy()} if c["opt_as_dict"] else np.column_stack((chi2, cube))
        np.random.seed(c["seed"])
        p0 = init_walkers(PS, distribution=c["distribution"], opt_data=opt, ntemps=c["ntemps"],
                          nwalkers=c["nwalkers"], fbest=c["fbest"])
        want = np.array([float.fromhex(v) for v in c["p0_hex"]]).reshape(c["shape"])
        assert p0.shape == tuple(c["shape"])
        assert np.array_equal(p0, want), c["distribution"]
        lo = np.array(PS.hypercube)[:, 0]; hi = np.array(PS.hypercube)[:, 1]
        if c["distribution"] in ("Uniform", "Gaussian"):
            assert ((p0 >= lo) & (p0 <= hi)).all()
        else:
            assert ((p0 >= 0.0) & (p0 <= 1.0)).all()
    capsys.readouterr()


def test_error_behaviour(golden, capsys):
    PS, chi2, cube = _inputs(golden)
    with pytest.raises(RuntimeError, match="nwalkers have to be >= 10"):
        init_walkers(PS, distribution="Uniform", ntemps=2, nwalkers=9)
    opt = np.column_stack((chi2, cube))
    with pytest.raises(ValueError, match="does not match with any"):
        init_walkers(PS, distribution="Cauchy", opt_data=opt, ntemps=2, nwalkers=10)
    with pytest.raises(AssertionError, match="fbest"):
        init_walkers(PS, distribution="Picked", opt_data=opt, ntemps=2, nwalkers=10, fbest=0.0)
    bad = opt.copy(); bad[0, 2] = 1.5
    with pytest.raises(AssertionError, match="cube"):
        init_walkers(PS, distribution="Picked", opt_data=bad, ntemps=2, nwalkers=10)
    assert init_walkers(PS, distribution="Gaussian", opt_data=None, ntemps=2, nwalkers=10) is None
    capsys.readouterr()
