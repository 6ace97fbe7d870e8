"""CPU: libttvb.so loads and exports every symbol include/ttvb.h declares; no compute calls."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT, has_gpu


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "ttvb.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ttvb_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from nauyaca_b200 import _lib
    L = _lib.lib()
    syms = _declared_symbols()
    assert set(syms) == set(_lib.EXPORTS)
    for s in syms:
        assert hasattr(L, s), s
    assert L.ttvb_version() == 1000


@pytest.mark.skipif(has_gpu(), reason="checks the no-GPU error path")
def test_no_cpu_fallback(systems):
    import nauyaca_b200 as nb
    with pytest.raises(nb.TTVBError) as e:
        nb.BatchEngine(systems["doc2"])
    assert "-2" in str(e.value) and "no CPU fallback" in str(e.value)


def test_create_rejects_bad_arguments():
    from nauyaca_b200 import _lib
    L = _lib.lib()
    h = C.c_void_p()
    assert L.ttvb_create(None, 0, C.byref(h)) == _lib.TTVB_E_INVALID
    s = _lib.ttvb_system()
    s.npla = 12
    assert L.ttvb_create(C.byref(s), 0, C.byref(h)) == _lib.TTVB_E_UNSUPPORTED
    s.npla, s.ndim, s.nconst = 2, 3, 3
    assert L.ttvb_create(C.byref(s), 0, C.byref(h)) == _lib.TTVB_E_INVALID
    assert b"7*npla" in L.ttvb_last_error()
    s.ndim, s.nconst, s.dt, s.ftime = 11, 3, 0.1, 10.0
    assert L.ttvb_create(C.byref(s), 0, C.byref(h)) == _lib.TTVB_E_INVALID      # mstar == 0
    assert b"mstar" in L.ttvb_last_error()


def test_sampler_rng_known_answers():
    """Philox4x32-10 against the Random123 known-answer vectors; the Feistel pairing really is
    a permutation (host code of libttvb, no GPU)."""
    import numpy as np
    from nauyaca_b200 import _lib
    L = _lib.lib()
    kat = [([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
           ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
           ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
            [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1])]
    for c, k, want in kat:
        out = (C.c_uint32 * 4)()
        assert L.ttvb_selftest_philox((C.c_uint32 * 4)(*c), (C.c_uint32 * 2)(*k), out) == 0
        assert list(out) == want
    for n in (1, 2, 22, 60, 1000, 4097):
        perm = np.zeros(n, np.uint32)
        assert L.ttvb_selftest_perm(C.c_uint32(n), C.c_uint32(9), C.c_uint32(n), perm.ctypes.data_as(C.c_void_p)) == 0
        assert sorted(perm.tolist()) == list(range(n))
    a = np.zeros(1000, np.uint32); b = np.zeros(1000, np.uint32)
    L.ttvb_selftest_perm(C.c_uint32(1000), C.c_uint32(1), C.c_uint32(2), a.ctypes.data_as(C.c_void_p))
    L.ttvb_selftest_perm(C.c_uint32(1000), C.c_uint32(1), C.c_uint32(3), b.ctypes.data_as(C.c_void_p))
    assert (a != b).sum() > 900                              # different keys, different pairings
