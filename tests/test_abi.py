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
