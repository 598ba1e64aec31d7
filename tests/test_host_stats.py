"""Host-side statistics of the product (mm_stats.h through the C ABI) against the oracle and the Boost vectors."""
import ctypes as C
import json
import os

import numpy as np

from metamaps_b200 import capi
from tests.conftest import GOLDEN, build_emu


def _lib():
    # the stat entry points are plain host code, identical in both builds; use the product library when present
    return capi.load(capi.LIB_PATH if os.path.exists(capi.LIB_PATH) else build_emu())


def test_min_hits_and_identity_match_oracle(oracle):
    lib = _lib()
    a = C.c_float(); b = C.c_float()
    for s in list(range(1, 400)) + [723, 940, 1500, 2500, 4700, 9000]:
        assert lib.mm_stat_min_hits_relaxed(s, 16, 80.0) == oracle.min_hits_relaxed(s, 16, 80.0), s
        for sh in {0, 1, 2, s // 50, s // 20, s // 10, s // 3, s}:
            lib.mm_stat_identity(sh, s, 16, C.byref(a), C.byref(b))
            assert (a.value, b.value) == oracle.identity(sh, s, 16), (s, sh)
    for pi in (70.0, 85.0, 90.0):
        for s in (10, 100, 1000):
            assert lib.mm_stat_min_hits_relaxed(s, 16, pi) == oracle.min_hits_relaxed(s, 16, pi)


def test_window_size_matches_oracle(oracle):
    lib = _lib()
    for (m, L) in ((1000, 2_025_370), (2000, 12_000_000_000), (1000, 26_762_276_280), (5000, 486_296), (2000, 100_000_000_000)):
        assert lib.mm_stat_recommended_window(1e-3, 16, 4, 80.0, m, L) == oracle.recommended_window(1e-3, 16, 80.0, m, L)
    # the example run of the reference: -m 2000 on miniSeq+H (26.76 GB FASTA) used w = 16
    assert lib.mm_stat_recommended_window(1e-3, 16, 4, 80.0, 2000, 26_762_276_280) == 16
