"""CPU, development container only: the UNMODIFIED reference nauyaca/utils.py, run over the
oracle-backed ttvfast shim, must agree with the oracle's own restatement of that plumbing
(bounds, cube_to_physical, constants, _ephemeris, _chi2, logL). Skipped where /root/reference
is absent (the GPU box)."""
import contextlib
import io
import os

import numpy as np
import pytest

pytestmark = pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="needs the reference checkout")


@pytest.fixture(scope="module")
def ref():
    from oracle import ref_plumbing
    with contextlib.redirect_stdout(io.StringIO()):
        nau, U = ref_plumbing.import_reference()
        PS = ref_plumbing.build_systems(nau)
    return U, PS


@pytest.mark.parametrize("name", ["doc2", "sys3"])
def test_reference_objectives_equal_oracle(ref, oracle, systems, name):
    U, PSs = ref
    PS = PSs[name]; osys = oracle.OracleSystem(systems[name])
    X = np.random.default_rng(4).random((40, PS.ndim))
    X[3, 1] = 1.5
    oc, ol, ost = osys.loglike(X, mode=0)
    for b, x in enumerate(X):
        with contextlib.redirect_stdout(io.StringIO()):
            ll = U.log_likelihood_func(x, PS)
            c2 = U.calculate_chi2(x, PS)
        # the shim divides the mass by Mearth_to_Msun and the oracle multiplies again: one
        # rounding can differ, so compare to 1e-9 relative (the north-star chi2 tolerance)
        if np.isinf(oc[b]):
            assert c2 == np.inf and ll == -np.inf
        else:
            assert abs(c2 - oc[b]) <= 1e-9 * oc[b] and abs(ll - ol[b]) <= 1e-9 * abs(ol[b]) + 1e-9


def test_reference_ephemeris_golden(ref, known):
    U, PSs = ref
    PS = PSs["sys3"]
    for g in ("G1", "G2"):
        with contextlib.redirect_stdout(io.StringIO()):
            eph = U.calculate_ephemeris(PS, known[g]["flat_params"])
        for pid, ge in known[g]["ephemeris"].items():
            assert [int(k) for k in eph[pid].keys()] == ge["epochs"]
            if ge["times"]:
                assert np.max(np.abs(np.array(list(eph[pid].values())) - np.array(ge["times"]))) <= 1e-7
    with contextlib.redirect_stdout(io.StringIO()):
        assert U.calculate_chi2_physical(known["G2"]["flat_params"], PS) == 3.6e21
        assert U.log_likelihood_func(known["G2"]["flat_params"], PS, cube=False) == -np.inf   # notebook cell 51


def test_opt_writer_matches_reference_writefile(ref, tmp_path):
    """Byte-for-byte the lines the reference's utils.writefile produces for the same text
    (optimizers.py:122-136, 160-166)."""
    U, PSs = ref
    from nauyaca_b200.optimizers import write_opt_files
    PS = PSs["doc2"]
    rows = [[45.09121274544064, list(np.linspace(0.05, 0.95, PS.ndim)), list(np.linspace(1.0, 300.0, PS.ndim))]]
    c, p = write_opt_files(PS, rows, path=str(tmp_path) + "/")
    rc, rp = str(tmp_path / "ref_cube.opt"), str(tmp_path / "ref_phys.opt")
    header = '#Chi2 ' + " ".join([i for i in PS.params_names.split()])
    for f in (rc, rp):
        U.writefile(f, "w", header, '%-15s' + ' %-11s' * (len(header.split()) - 1) + '\n')
    f2, x2_cube, x2_phys = rows[0]
    info = '{} \t {} \n'.format(str(f2), "  ".join([str(it) for it in x2_cube]))
    U.writefile(rc, 'a', info, '%-30s' + ' %-11s' * (len(info.split()) - 1) + '\n')
    info = ' ' + str(f2) + ' ' + "  ".join([str(it) for it in x2_phys])
    U.writefile(rp, 'a', info, '%-30s' + ' %-11s' * (len(info.split()) - 1) + '\n')
    assert open(c, "rb").read() == open(rc, "rb").read()
    assert open(p, "rb").read() == open(rp, "rb").read()
