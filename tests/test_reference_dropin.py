"""Drop-in proof with the reference's OWN classes (north star: "keeps the PlanetarySystem /
Optimizers / MCMC API, so the change is a drop-in behind those calls").

The unmodified reference package is imported from baseline/_ref (written by
__graft_entry__.build() in the development container; it travels to the GPU box) or from the
checkout. CPU part: `system_spec` of a REAL PlanetarySystem equals the frozen golden systems.
GPU part: the reference's own utils functions run over the GPU-backed `ttvfast` module and give the
oracle's values; the mirror functions accept the real PSystem object; ptemcee-shaped `pool.map`
with the reference's likelihood object is one GPU batch; the reference's MCMC.run runs unmodified
on top of GpuPool + the restated sampler."""
import contextlib
import io
import os
import sys

import numpy as np
import pytest

from nauyaca_b200 import compat  # noqa: E402
from oracle import ref_systems as _ref_systems  # noqa: E402  (test infrastructure: builds the notebook systems)

pytestmark = pytest.mark.skipif(compat.find_reference() is None, reason="reference package not available "
                                "(baseline/_ref is written by __graft_entry__.build() in the dev container)")

HOT_KEYS = ["mstar", "rstarAU", "t0", "ftime", "dt", "npla", "ndim", "planets_IDs", "bounds", "bi", "bf", "hypercube",
            "constant_params", "observed", "second_term_sum"]


def test_system_spec_of_real_planetary_system(systems):
    """a12: engine.system_spec reads a real nauyaca.PlanetarySystem (planetarysystem.py:184-350)."""
    if os.path.isdir("/root/reference"):
        from oracle import ref_plumbing                   # same session as test_reference_plumbing: share its import
        with contextlib.redirect_stdout(io.StringIO()):
            nau, _U = ref_plumbing.import_reference()
    else:
        nau = compat.import_reference()
    from nauyaca_b200.engine import system_spec
    PS = _ref_systems.build(nau, compat.reference_inputs())
    for name in ("doc2", "sys3", "sys3f"):
        assert type(PS[name]).__module__.startswith("nauyaca.")
        spec = system_spec(PS[name])
        for k in HOT_KEYS:
            assert spec[k] == systems[name][k], (name, k)


# ------------------------------------------------------------------------------------------ GPU
@pytest.fixture(scope="module")
def ref():
    with contextlib.redirect_stdout(io.StringIO()):
        nau = compat.import_reference()
        PS = _ref_systems.build(nau, compat.reference_inputs())
    import nauyaca.utils as RU
    assert RU.ttvfast.ttvfast is compat.ttvfast, "the reference must be running over the GPU-backed ttvfast module"
    return nau, RU, PS


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["doc2", "sys3"])
def test_reference_objectives_over_gpu_ttvfast_equal_oracle(ref, oracle, systems, name):
    """Seam S1: utils.log_likelihood_func / calculate_chi2 of the REFERENCE, ttvfast = ttvb_ephemeris."""
    nau, RU, PSs = ref
    PS = PSs[name]; osys = oracle.OracleSystem(systems[name])
    X = np.random.default_rng(4).random((24, PS.ndim))
    X[3, 1] = 1.5
    oc, ol, _ost = osys.loglike(X, mode=0)
    same_bits = 0
    for b, x in enumerate(X):
        with contextlib.redirect_stdout(io.StringIO()):
            ll = RU.log_likelihood_func(x, PS)
            c2 = RU.calculate_chi2(x, PS)
        if np.isinf(oc[b]):
            assert c2 == np.inf and ll == -np.inf
        else:
            assert abs(c2 - oc[b]) <= 1e-9 * oc[b] and abs(ll - ol[b]) <= 1e-9 * abs(ol[b]) + 1e-9
            same_bits += (c2 == oc[b])
    assert same_bits >= 20          # the Msun detour of utils.py:62 is undone exactly (compat._mass_in_mearth)


@pytest.mark.gpu
def test_reference_ephemeris_golden_over_gpu_ttvfast(ref, known):
    nau, RU, PSs = ref
    PS = PSs["sys3"]
    for g in ("G1", "G2"):
        with contextlib.redirect_stdout(io.StringIO()):
            eph = RU.calculate_ephemeris(PS, known[g]["flat_params"])
        for pid, ge in known[g]["ephemeris"].items():
            assert [int(k) for k in eph[pid].keys()] == ge["epochs"]
            if ge["times"]:
                assert np.max(np.abs(np.array(list(eph[pid].values())) - np.array(ge["times"]))) <= 1e-7
    with contextlib.redirect_stdout(io.StringIO()):
        assert RU.calculate_chi2_physical(known["G2"]["flat_params"], PS) == 3.6e21
        assert RU.log_likelihood_func(known["G2"]["flat_params"], PS, cube=False) == -np.inf


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["doc2", "sys3", "sys3f"])
def test_mirror_functions_take_the_real_planetary_system(ref, oracle, systems, name):
    """Seam S2: nauyaca_b200.utils.* called with the reference's PSystem object give what the
    reference's own functions give (and the oracle, bit for bit)."""
    import nauyaca_b200.utils as GU
    nau, RU, PSs = ref
    PS = PSs[name]; osys = oracle.OracleSystem(systems[name])
    X = np.random.default_rng(9).random((16, PS.ndim))
    oc, ol, _ = osys.loglike(X, mode=0)
    assert np.array_equal(GU.calculate_chi2_batch(X, PS), oc)
    assert np.array_equal(GU.log_likelihood_batch(X, PS), ol)
    for x in X[:4]:
        with contextlib.redirect_stdout(io.StringIO()):
            r_flat = RU.cube_to_physical(PS, x)
            r_ll = RU.log_likelihood_func(x, PS)
            r_ind = RU.calculate_chi2_physical(r_flat, PS, individual=True)
        assert list(GU.cube_to_physical(PS, x)) == list(r_flat)
        assert abs(GU.log_likelihood_func(x, PS) - r_ll) <= 1e-9 * abs(r_ll)
        g_ind = GU.calculate_chi2_physical(r_flat, PS, individual=True)
        assert g_ind.keys() == r_ind.keys()
        for k in r_ind:
            assert abs(g_ind[k] - r_ind[k]) <= 1e-9 * r_ind[k]
        assert GU.intervals(PS.hypercube, x) == RU.intervals(PS.hypercube, x)


@pytest.mark.gpu
def test_gpu_pool_batches_the_reference_likelihood_object(ref, oracle, systems):
    """Seam S3: ptemcee calls pool.map(LikePriorEvaluator, vectors) with logl = the reference's
    log_likelihood_func and logp = MCMC.logprior (mcmc.py:195-211): one GPU launch for the block."""
    import nauyaca_b200 as nb
    from nauyaca_b200.ptsampler import LikePriorEvaluator
    nau, RU, PSs = ref
    PS = PSs["doc2"]; osys = oracle.OracleSystem(systems["doc2"])
    ev = LikePriorEvaluator(logl=RU.log_likelihood_func, logp=nau.MCMC.logprior, loglargs=(PS, True, False),
                            logpkwargs={'psystem': PS})
    X = np.random.default_rng(2).random((500, PS.ndim))
    X[7, 0] = -0.1
    eng = nb.utils.engine_for(PS)
    n0 = eng.launch_count()
    out = nb.GpuPool().map(ev, list(X))
    assert eng.launch_count() - n0 == 1
    _oc, ol, _ = osys.loglike(X, mode=0)
    assert [o[0] for o in out] == list(ol) and all(o[1] == 0.0 for o in out)


@pytest.mark.gpu
def test_reference_mcmc_run_unmodified(ref, tmp_path, systems):
    """The reference's MCMC.run (mcmc.py:97-312), not a line changed, drives the GPU path:
    ptemcee -> nauyaca_b200.ptsampler, Pool -> GpuPool, h5py -> compat.H5Lite (h5py is absent)."""
    import nauyaca_b200 as nb
    nau, RU, PSs = ref
    PS = PSs["doc2"]
    rng = np.random.default_rng(1)
    ntemps, nwalkers = 3, 24
    p0 = 0.3 + 0.4 * rng.random((ntemps, nwalkers, PS.ndim))
    eng = nb.utils.engine_for(PS)
    n0 = eng.launch_count()
    with contextlib.redirect_stdout(io.StringIO()) as log:
        mc = nau.MCMC(PS, p0=p0, itmax=6, intra_steps=2, tmax=50.0, cores=2, path=str(tmp_path) + "/", suffix="_t",
                      verbose=True)
        sampler = mc.run()
    # 1 launch for p0 + 2 half-steps per iteration, each ONE batch of ntemps*nwalkers/2 vectors
    assert eng.launch_count() - n0 == 1 + 2 * 6
    assert type(sampler).__module__ == "nauyaca_b200.ptsampler"
    assert "Maximum number of iterations reached" in log.getvalue() or "Iterations performed" in log.getvalue()
    import h5py
    with h5py.File(str(tmp_path / "SystemX_t.hdf5"), "r") as f:
        chains = f["CHAINS"][()]
        assert chains.shape == (ntemps, nwalkers, 3, PS.ndim) and f["INDEX"][()][0] == 2
        assert np.isfinite(f["MAP"][()]).all() and ((chains >= 0) & (chains <= 1)).all()
    assert os.path.exists(str(tmp_path / "SystemX_t.best"))
