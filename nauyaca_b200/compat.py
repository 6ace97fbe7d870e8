"""Running the UNMODIFIED reference package (nauyaca) on top of the GPU path.

nauyaca reaches its engine through ``ttvfast.ttvfast(...)`` (utils.py:72-74), its sampler through
``ptemcee.Sampler(pool=Pool(cores))`` (mcmc.py:193-211) and its chain file through ``h5py``
(mcmc.py:329-401). None of the three is installable next to this package (no network), and the
first two are exactly the seams the GPU path replaces. This module provides, as importable
stand-ins under the names the reference imports:

``ttvfast``   seam S1: ``ttvfast.ttvfast(planets, stellar_mass, time, dt, total, rv_times, input_flag)``
              and ``ttvfast.models.Planet`` over ``ttvb_ephemeris`` (one launch per call; same
              return layout: five lists of 5000 padded with -2).
``ptemcee``   = ``nauyaca_b200.ptsampler`` (the restated PT sampler; same constructor/sample()).
``h5py``      ``H5Lite``: an h5py-SHAPED container (File / create_dataset / slicing) stored as
              .npz. It is NOT the HDF5 format; it exists so that ``MCMC.run`` and
              ``utils.extract_best_solutions`` run unmodified where h5py is missing.

``import_reference()`` installs whichever of these are missing into ``sys.modules``, stubs the
plotting imports, imports ``nauyaca`` from ``path`` and points ``nauyaca.mcmc.Pool`` at
``GpuPool`` (the one-line change of INTEGRATION.md, seam S3). Nothing of the reference is
copied or modified; there is still no CPU fallback underneath.
"""
import os
import sys
import types

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_MEARTH = 3.0034893488507934e-06        # reference constants.py:65
N_EVENTS = 5000                          # slots of ttvfast-python's event table (utils.py:76-79)


# ------------------------------------------------------------------------------------ ttvfast
class Planet:
    """ttvfast.models.Planet as constructed by utils.py:60-70 (mass in Msun, angles in degrees)."""

    def __init__(self, mass, period, eccentricity, inclination, longnode, argument, mean_anomaly):
        self.mass, self.period, self.eccentricity = mass, period, eccentricity
        self.inclination, self.longnode, self.argument, self.mean_anomaly = inclination, longnode, argument, mean_anomaly


def _mass_in_mearth(mass_msun):
    """The kernel takes masses in Mearth and multiplies by Mearth_to_Msun itself, as utils.py:62
    does. Return the Mearth value whose product with the constant is `mass_msun` again (the
    quotient or one of its neighbours), so that the detour through Msun changes no bit."""
    m = mass_msun / _MEARTH
    for cand in (m, np.nextafter(m, np.inf), np.nextafter(m, -np.inf)):
        if cand * _MEARTH == mass_msun:
            return float(cand)
    return float(m)


def ttvfast(planets, stellar_mass, time, dt, total, rv_times=None, input_flag=1):
    """``ttvfast.ttvfast`` for the call nauyaca makes (input_flag=1, no RVs) on the GPU."""
    from . import utils as _u
    if input_flag != 1 or rv_times is not None:
        raise NotImplementedError("covers the call nauyaca makes: input_flag=1, rv_times=None")
    flat = []
    for p in planets:
        flat += [_mass_in_mearth(p.mass), p.period, p.eccentricity, p.inclination, p.argument, p.mean_anomaly,
                 p.longnode]
    eng = _u._raw_engine(len(planets), stellar_mass, time, total, dt)
    n, pl, ep, tm, rs, vs, _st = eng.ephemeris(np.asarray(flat, dtype=np.float64)[None, :], mode=2, cap=N_EVENTS)
    return {"positions": [pl[0].tolist(), ep[0].tolist(), tm[0].tolist(), rs[0].tolist(), vs[0].tolist()],
            "rv": None}


def ttvfast_module():
    """A module object that can stand in for ``ttvfast`` (``sys.modules['ttvfast']``)."""
    m = types.ModuleType("ttvfast")
    m.__path__ = []
    m.ttvfast = ttvfast
    models = types.ModuleType("ttvfast.models")
    models.Planet = Planet
    m.models = models
    return m, models


# ------------------------------------------------------------------------------------ h5py stand-in
class _Dataset:
    def __init__(self, arr):
        self.a = arr

    def __getitem__(self, k):
        v = self.a[k]
        return v.item() if (self.a.dtype.kind in "US" and np.ndim(v) == 0) else v

    def __setitem__(self, k, v):
        self.a[k] = v

    shape = property(lambda self: self.a.shape)
    dtype = property(lambda self: self.a.dtype)

    def __len__(self):
        return len(self.a)

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self.a, dtype=dtype)


class H5LiteFile:
    """``h5py.File(name, mode)`` look-alike over a .npz payload (modes 'w', 'r', 'r+', 'a')."""

    def __init__(self, name, mode="r", **_kw):
        self.filename, self.mode = name, mode
        self._d = {}
        if mode in ("r", "r+", "a") and os.path.exists(name):
            with np.load(name, allow_pickle=False) as z:
                self._d = {k: _Dataset(np.array(z[k])) for k in z.files}
        elif mode in ("r", "r+"):
            raise OSError(f"Unable to open file (unable to open file: name = '{name}')")

    def create_dataset(self, name, shape=None, dtype=None, data=None, **_kw):
        if data is not None:
            arr = np.array(data, dtype=dtype)
        else:
            arr = np.zeros(shape, dtype=dtype or "f8")
        self._d[name] = _Dataset(arr)
        return self._d[name]

    def __setitem__(self, name, value):
        self._d[name] = _Dataset(np.array(value))

    def __getitem__(self, name):
        return self._d[name]

    def __contains__(self, name):
        return name in self._d

    def keys(self):
        return self._d.keys()

    def flush(self):
        if self.mode != "r":
            with open(self.filename, "wb") as fh:
                np.savez(fh, **{k: v.a for k, v in self._d.items()})

    def close(self):
        self.flush()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
        return False


def h5lite_module():
    m = types.ModuleType("h5py")
    m.File = H5LiteFile
    m.__h5lite__ = True
    return m


# ------------------------------------------------------------------------------------ reference import
def find_reference(path=None):
    """Directory that contains the ``nauyaca`` package: `path`, $NAUYACA_REF, the in-tree install
    ``baseline/_ref`` (written by ``__graft_entry__.build()``, git-ignored) or the reference
    checkout of the development container."""
    root = os.path.dirname(_HERE)
    for cand in (path, os.environ.get("NAUYACA_REF"), os.path.join(root, "baseline", "_ref"), "/root/reference"):
        if cand and os.path.isfile(os.path.join(cand, "nauyaca", "__init__.py")):
            return cand
    return None


def _importable(name):
    import importlib.util
    try:
        return importlib.util.find_spec(name) is not None
    except (ImportError, ValueError):
        return False


def import_reference(path=None, gpu_pool=True):
    """Import the unmodified reference package with its missing third-party modules replaced as
    described in the module docstring. Returns the ``nauyaca`` module.
    Raises ImportError if the package is nowhere to be found."""
    ref = find_reference(path)
    if ref is None:
        raise ImportError("the reference package nauyaca was not found (path, $NAUYACA_REF, baseline/_ref)")
    if "ttvfast" not in sys.modules and not _importable("ttvfast"):
        sys.modules["ttvfast"], sys.modules["ttvfast.models"] = ttvfast_module()
    if "ptemcee" not in sys.modules and not _importable("ptemcee"):
        from . import ptsampler
        sys.modules["ptemcee"] = ptsampler
    if "h5py" not in sys.modules and not _importable("h5py"):
        sys.modules["h5py"] = h5lite_module()
    # plotting / regression imports of plots.py and utils.py: not on the path, stubbed if absent
    for n in ["seaborn", "corner", "matplotlib", "matplotlib.pyplot", "matplotlib.ticker", "matplotlib.gridspec",
              "mpl_toolkits", "mpl_toolkits.axes_grid1", "sklearn", "sklearn.linear_model"]:
        if n not in sys.modules and not _importable(n):
            m = types.ModuleType(n)
            m.__path__ = []
            sys.modules[n] = m
    import importlib
    for n, attr in (("mpl_toolkits.axes_grid1", "make_axes_locatable"), ("sklearn.linear_model", "LinearRegression"),
                    ("matplotlib.ticker", "FormatStrFormatter")):
        mod = sys.modules.get(n) or importlib.import_module(n)       # installed on this machine: the real one
        if not hasattr(mod, attr):
            setattr(mod, attr, object)
    if ref not in sys.path:
        sys.path.insert(0, ref)
    import nauyaca
    if gpu_pool:
        from .pool import GpuPool
        import nauyaca.mcmc as rmcmc
        rmcmc.Pool = lambda processes=None, **kw: GpuPool()        # mcmc.py:193, seam S3
    return nauyaca


def reference_inputs(ref=None):
    """Directory with the reference's observation tables (Documentation/inputs, Examples/inputs):
    the checkout itself, or ``baseline/_ref/inputs`` on the GPU box."""
    ref = find_reference(ref)
    if ref is None:
        return None
    for cand in (os.path.join(ref, "inputs"), ref):
        if os.path.isdir(os.path.join(cand, "Documentation", "inputs")):
            return cand
    return None
