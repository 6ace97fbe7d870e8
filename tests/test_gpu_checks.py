"""GPU: the self-checking kernel build (-DTTVB_CHECKS) on the cases a compute-sanitizer run would
cover - time-segmented launches with the inter-CTA hand-over, a saturated event queue, table mode, the
device sampler. compute-sanitizer is closed on this GPU pool (profiles/r2a_compute_sanitizer_closed.txt),
so the kernels check their own queue / ring / table indices and bound their waits; zero violations and
bit-equal results are required."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("nseg", ["3", "1"])
def test_self_checking_build_reports_no_violation(nseg):
    from nauyaca_b200 import _lib
    lib = _lib.build_checks()
    env = dict(os.environ, TTVB_LIB=lib, TTVB_NSEG=nseg)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "checks_case.py")], env=env, capture_output=True,
                         text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    rep = json.loads(out.stdout.strip().splitlines()[-1])
    assert rep["checks_case"] == "ok" and rep["check_rc"] == 0 and rep["lib"] == "libttvb_checks.so"
    assert rep["violations"] == [0] * 8, rep


def test_product_build_has_no_checks():
    import ctypes
    from nauyaca_b200 import _lib
    counts = (ctypes.c_int32 * 8)()
    assert _lib.lib().ttvb_check_failures(0, counts) == _lib.TTVB_E_UNSUPPORTED and list(counts) == [-1] * 8
