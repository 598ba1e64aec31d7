"""The C-ABI library exports exactly what include/metamaps_b200.h declares, and refuses to run without a GPU."""
import ctypes
import os
import re

import pytest

from metamaps_b200 import capi
from tests.conftest import ROOT


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "metamaps_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mm_[a-z0-9_]+)\s*\(", txt)))


def test_header_and_binding_agree():
    assert header_symbols() == sorted(capi.SYMBOLS.keys())


def test_library_exports_every_declared_symbol():
    assert os.path.exists(capi.LIB_PATH), "build the CUDA library first (__graft_entry__.build())"
    lib = ctypes.CDLL(capi.LIB_PATH)
    for name in header_symbols():
        assert hasattr(lib, name), name
    lib.mm_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.mm_version()


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = capi.load()
    h = ctypes.c_void_p()
    rc = lib.mm_ctx_create(0, ctypes.byref(h))
    assert rc == -19 and b"no CUDA device" in lib.mm_last_error()     # MM_ENODEV: fails loudly, no CPU path
    with pytest.raises(capi.MMError):
        capi.Context(0, lib)


def test_missing_library_is_an_error(tmp_path):
    with pytest.raises(FileNotFoundError):
        capi.load(str(tmp_path / "nope.so"))
