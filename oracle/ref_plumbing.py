#!/usr/bin/env python
"""Run the UNMODIFIED reference plumbing (/root/reference/nauyaca/utils.py) over the oracle-backed
`ttvfast` shim (TEST INFRASTRUCTURE; development container only: /root/reference does not
travel to the GPU box).

  python oracle/ref_plumbing.py            # reference-plumbing CPU baseline, Pool(os.cpu_count())

Also imported by tests/test_reference_plumbing.py to check that the reference's own
log_likelihood_func / calculate_chi2 / calculate_ephemeris give what the oracle (and hence
the GPU path) gives."""
import os
import sys
import time
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def import_reference():
    """import nauyaca from the read-only checkout with the absent third-party modules stubbed
    and `ttvfast` pointing at the shim. Returns (nauyaca, nauyaca.utils)."""
    if not os.path.isdir(REF):
        raise RuntimeError("/root/reference is not available on this machine")
    sys.path.insert(0, os.path.join(HERE, "ttvfast_shim"))
    for n in ["h5py", "ptemcee", "seaborn", "corner", "matplotlib", "matplotlib.pyplot", "matplotlib.ticker",
              "matplotlib.gridspec", "mpl_toolkits", "mpl_toolkits.axes_grid1", "sklearn", "sklearn.linear_model"]:
        if n not in sys.modules:
            m = types.ModuleType(n)
            m.__path__ = []
            sys.modules[n] = m
    sys.modules["mpl_toolkits.axes_grid1"].make_axes_locatable = lambda *a, **k: None
    sys.modules["sklearn.linear_model"].LinearRegression = object
    sys.modules["matplotlib.ticker"].FormatStrFormatter = object
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import nauyaca
    from nauyaca import utils
    return nauyaca, utils


def build_systems(nau):
    d = os.path.join(REF, "Documentation", "inputs")
    Pb = nau.SetPlanet("Planet-b"); Pb.load_ttvs(os.path.join(d, "planetb_ttvs.dat"))
    Pb.mass = [1, 50.]; Pb.period = [5.35, 5.36]; Pb.ecc = [0., 0.2]; Pb.inclination = [90, 90]; Pb.ascending_node = [88.36, 88.36]
    Pc = nau.SetPlanet("Planet-c"); Pc.load_ttvs(os.path.join(d, "planetc_ttvs.dat"))
    Pc.mass = [1, 50]; Pc.period = [11.83, 11.84]; Pc.ecc = [0, .2]; Pc.inclination = [90, 90]; Pc.ascending_node = [70, 110.]
    doc2 = nau.PlanetarySystem("SystemX", mstar=0.9132, rstar=0.8632); doc2.add_planets([Pb, Pc])
    doc2.simulation(t0=0.0, ftime=221, dt=0.18)
    e = os.path.join(REF, "Examples", "inputs")
    P1 = nau.SetPlanet('Planet-b'); P1.mass = [0.001, 10]; P1.period = [10.47, 10.48]; P1.ecc = [0.0, 0.1]
    P1.inclination = [90, 90]; P1.ascending_node = [60, 120]; P1.load_ttvs(os.path.join(e, "3pl_planet1_ttvs.dat"))
    P2 = nau.SetPlanet('Planet-c'); P2.mass = [0.001, 10]; P2.period = [14.10, 14.11]; P2.ecc = [0.0, 0.1]
    P2.inclination = [85, 95]; P2.ascending_node = [60, 120]; P2.load_ttvs(os.path.join(e, "3pl_planet2_ttvs.dat"))
    P3 = nau.SetPlanet('Planet-d'); P3.mass = [0.001, 10]; P3.period = [23.57, 23.58]; P3.ecc = [0.0, 0.1]
    P3.inclination = [90.3, 90.3]; P3.ascending_node = [88.5, 88.5]; P3.load_ttvs(os.path.join(e, "3pl_planet3_ttvs.dat"))
    sys3 = nau.PlanetarySystem("MySystem", mstar=0.522, rstar=0.4422); sys3.add_planets([P1, P2, P3])
    sys3.simulation(t0=0, ftime=500)
    return {"doc2": doc2, "sys3": sys3}


_PS = None
_U = None


def _init(name):
    global _PS, _U
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        nau, _U = import_reference()
        _PS = build_systems(nau)[name]


def _work(x):
    return _U.log_likelihood_func(x, _PS)


def main():
    import multiprocessing as mp
    import numpy as np
    cores = os.cpu_count() or 1
    for name, n in (("doc2", 4000), ("sys3", 4000)):
        _init(name)
        X = np.random.default_rng(20240607).random((n, _PS.ndim))
        t = time.perf_counter(); [_work(x) for x in X[:200]]; t1 = (time.perf_counter() - t) / 200
        with mp.get_context("fork").Pool(cores, initializer=_init, initargs=(name,)) as pool:
            pool.map(_work, list(X[:cores * 4]))
            t = time.perf_counter(); pool.map(_work, list(X)); wall = time.perf_counter() - t
        print(f"{name}: reference utils.log_likelihood_func over the ttvfast shim: {1 / t1:8.0f} evals/s on 1 core, "
              f"{n / wall:8.0f} evals/s with Pool({cores}).map  ({1e3 * t1:.3f} ms per call)")


if __name__ == "__main__":
    main()
