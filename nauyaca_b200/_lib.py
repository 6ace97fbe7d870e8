"""ctypes binding of libttvb.so (C ABI in include/ttvb.h) and its in-tree build.

There is no CPU fallback: if the library is missing or no CUDA device is present, calls
raise. The library is built in-tree by ``build()`` (nvcc, sm_100a) so that it travels with
the repository snapshot.
"""
import ctypes as C
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TTVB_LIB") or os.path.join(_HERE, "libttvb.so")   # TTVB_LIB: kernel experiments
SRC = os.path.join(_HERE, "csrc", "ttvb.cu")
DEPS = [SRC, os.path.join(_HERE, "csrc", "ttvb_kernels.cuh"), os.path.join(_HERE, "csrc", "ttvb_math.cuh"),
        os.path.join(_HERE, "csrc", "ttvb_pt.cuh"),
        os.path.join(os.path.dirname(_HERE), "include", "ttvb.h")]

# -fmad=false: only the fma() calls written in the source are fused, so the device
# arithmetic is the same IEEE-754 operation sequence as the CPU oracle's.
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-fmad=false",
              "-std=c++17", "-Xcompiler", "-fPIC", "-shared"]

TTVB_OK, TTVB_E_INVALID, TTVB_E_NO_DEVICE, TTVB_E_CUDA, TTVB_E_UNSUPPORTED = 0, -1, -2, -3, -4
MODE_CUBE, MODE_PHYSICAL, MODE_PHYSICAL_FULL = 0, 1, 2
ST_OUT_OF_BOUNDS, ST_KEPLER_FAIL, ST_BAD_TRANSIT, ST_EVENT_OVERFLOW, ST_INVALID_INPUT = 1, 2, 4, 8, 16
EVENT_CAP = 5000
MAX_PLANETS = 9

EXPORTS = ["ttvb_last_error", "ttvb_version", "ttvb_device_count", "ttvb_create", "ttvb_destroy",
           "ttvb_set_cube_bounds", "ttvb_nsteps", "ttvb_loglike", "ttvb_ephemeris", "ttvb_cube_to_physical",
           "ttvb_last_kernel_ms", "ttvb_launch_count", "ttvb_stream_retries", "ttvb_fp64_peak",
           "ttvb_pt_create", "ttvb_pt_destroy", "ttvb_pt_set_walkers", "ttvb_pt_run", "ttvb_pt_get", "ttvb_pt_half_begin", "ttvb_pt_half_end", "ttvb_pt_proposal_logl",
           "ttvb_selftest_philox", "ttvb_selftest_perm", "ttvb_selftest_seed_scan", "ttvb_check_failures"]


class TTVBError(RuntimeError):
    pass


class ttvb_system(C.Structure):
    _fields_ = [("npla", C.c_int32), ("ndim", C.c_int32), ("nconst", C.c_int32), ("nobs", C.c_int32),
                ("mstar", C.c_double), ("t0", C.c_double), ("ftime", C.c_double), ("dt", C.c_double),
                ("rstar_au", C.c_double), ("second_term", C.c_double),
                ("bi", C.c_void_p), ("bf", C.c_void_p),
                ("cube_lo", C.c_void_p), ("cube_hi", C.c_void_p),
                ("phys_lo", C.c_void_p), ("phys_hi", C.c_void_p),
                ("const_idx", C.c_void_p), ("const_val", C.c_void_p),
                ("obs_planet", C.c_void_p), ("obs_epoch", C.c_void_p),
                ("obs_time", C.c_void_p), ("obs_sigma", C.c_void_p)]


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in DEPS)


CHECKS_LIB_PATH = os.path.join(_HERE, "libttvb_checks.so")


def build_checks(force=False):
    """The self-checking build (nvcc -DTTVB_CHECKS, 2- and 3-planet kernels only): see
    ttvb_check_failures in include/ttvb.h. Used by tests/test_gpu_checks.py in a subprocess
    (TTVB_LIB=...); never loaded by the product path."""
    if not force and os.path.exists(CHECKS_LIB_PATH) and \
            all(os.path.getmtime(d) <= os.path.getmtime(CHECKS_LIB_PATH) for d in DEPS if os.path.exists(d)):
        return CHECKS_LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    tmp = CHECKS_LIB_PATH + f".tmp{os.getpid()}"
    subprocess.check_call([nvcc] + NVCC_FLAGS + ["-DTTVB_CHECKS", "-DTTVB_DEV_ONLY_23", "-o", tmp, SRC])
    os.replace(tmp, CHECKS_LIB_PATH)
    return CHECKS_LIB_PATH


def build(force=False, verbose=False):
    """Compile libttvb.so for sm_100a with nvcc (cross-compiles without a GPU)."""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    tmp = LIB_PATH + f".tmp{os.getpid()}"
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp, SRC]
    subprocess.check_call(cmd)
    os.replace(tmp, LIB_PATH)          # readers never see a half-written library
    return LIB_PATH


def ensure_built(builder=True, timeout_s=600.0):
    """Build libttvb.so if it is missing or older than its sources (entry points that may run on a
    fresh checkout call this: the test session, bench.py, smoke()). Not a fallback: what gets
    built is the CUDA library, and without nvcc this raises. builder=False (ranks other than the
    first of a multi-process launch): wait for the builder instead of building."""
    if "TTVB_LIB" in os.environ or not needs_build():
        return LIB_PATH
    if builder:
        build(force=True)
        return LIB_PATH
    import time
    t0 = time.time()
    while needs_build():
        if time.time() - t0 > timeout_s:
            raise TTVBError(f"{LIB_PATH} was not built within {timeout_s:.0f} s")
        time.sleep(0.5)
    return LIB_PATH


_lib = None


def lib():
    """Load libttvb.so; raises TTVBError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TTVBError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(nauyaca_b200 has no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    L.ttvb_last_error.restype = C.c_char_p
    L.ttvb_create.argtypes = [C.POINTER(ttvb_system), C.c_int, C.POINTER(C.c_void_p)]
    L.ttvb_destroy.argtypes = [C.c_void_p]
    L.ttvb_set_cube_bounds.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.ttvb_nsteps.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
    L.ttvb_loglike.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p,
                               C.c_void_p, C.c_void_p, C.c_void_p]
    L.ttvb_ephemeris.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int32, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_void_p]
    L.ttvb_cube_to_physical.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    L.ttvb_last_kernel_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
    L.ttvb_launch_count.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
    L.ttvb_stream_retries.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
    L.ttvb_pt_create.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_double, C.c_double, C.c_double,
                                 C.c_uint64, C.c_int, C.POINTER(C.c_void_p)]
    L.ttvb_pt_destroy.argtypes = [C.c_void_p]
    L.ttvb_pt_set_walkers.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.ttvb_pt_run.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_void_p]
    L.ttvb_pt_get.argtypes = [C.c_void_p] + [C.c_void_p] * 8
    L.ttvb_pt_half_begin.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_void_p]
    L.ttvb_pt_half_end.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    L.ttvb_pt_proposal_logl.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.ttvb_fp64_peak.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_float)]
    L.ttvb_selftest_seed_scan.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.ttvb_check_failures.argtypes = [C.c_int, C.c_void_p]
    _lib = L
    return L


def check(rc):
    if rc != TTVB_OK:
        raise TTVBError(f"libttvb error {rc}: {lib().ttvb_last_error().decode()}")
