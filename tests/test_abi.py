"""The C-ABI library loads without a GPU, exports every symbol include/blokus_b200.h declares, and
fails loudly (no CPU fallback) when no CUDA device is present."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "blokus_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(bk_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported():
    from blokus_self_play import DEFAULT_LIB
    assert os.path.exists(DEFAULT_LIB), "build the library first: python __graft_entry__.py"
    dll = C.CDLL(DEFAULT_LIB)
    syms = header_symbols()
    assert len(syms) >= 40
    missing = [s for s in syms if not hasattr(dll, s)]
    assert not missing, missing


def test_binding_covers_header():
    from blokus_self_play import _lib
    assert sorted(_lib._SIGNATURES) == header_symbols()


def test_no_cpu_fallback_without_device():
    from blokus_self_play import Lib, DEFAULT_LIB, GameBatch
    lib = Lib(DEFAULT_LIB)
    if lib.bk_device_count() > 0:
        pytest.skip("a CUDA device is visible; the no-device behaviour is checked on CPU boxes")
    with pytest.raises(RuntimeError, match="no CUDA device"):
        GameBatch(1, lib=lib)
    h = C.c_void_p()
    rc = lib.bk_env_create(1, 0, C.byref(h))
    assert rc == -2 and b"no CPU fallback" in lib.bk_last_error()


def test_missing_library_is_loud(tmp_path):
    from blokus_self_play import Lib
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Lib(str(tmp_path / "nope.so"))
