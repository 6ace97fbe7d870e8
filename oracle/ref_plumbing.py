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


def reference_dir():
    """The checkout (development container) or the copy __graft_entry__.build() left in
    baseline/_ref (travels to the GPU box); None if neither exists."""
    for cand in (REF, os.path.join(os.path.dirname(HERE), "baseline", "_ref")):
        if os.path.isfile(os.path.join(cand, "nauyaca", "__init__.py")):
            return cand
    return None


def import_reference():
    """import nauyaca (unmodified) with the absent third-party modules stubbed and `ttvfast`
    pointing at the oracle-backed shim. Returns (nauyaca, nauyaca.utils)."""
    REF = reference_dir()
    if REF is None:
        raise RuntimeError("the reference package is not available on this machine")
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


def inputs_dir():
    ref = reference_dir()
    for cand in (os.path.join(ref, "inputs"), ref):
        if os.path.isdir(os.path.join(cand, "Documentation", "inputs")):
            return cand
    raise RuntimeError("the reference's observation tables were not found")


def build_systems(nau):
    """doc2 / sys3 / sys3f built with the reference's own classes (oracle/ref_systems.py)."""
    root = os.path.dirname(HERE)
    if root not in sys.path:
        sys.path.insert(0, root)
    from oracle import ref_systems
    return ref_systems.build(nau, inputs_dir())


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
    os.environ["TTVO_NATIVE"] = "1"     # timing: host divide/sqrt, not the emulated GPU seeds
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
