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


def test_identity_batch_matches_numpy_and_printf():
    """mm_stat_identity_batch = computeMap.hpp:403-408 in float + the 6-significant-digit text round trip."""
    import numpy as np
    from metamaps_b200 import capi, pipeline
    from tests.conftest import build_emu
    lib = capi.load(build_emu())
    rng = np.random.default_rng(3)
    s = rng.integers(1, 3000, 20000).astype(np.int32)
    sh = (rng.random(20000) * (s + 1)).astype(np.int32)
    sh[:50] = 0; sh[50:100] = s[50:100]
    a, b = capi.identity_batch(lib, sh, s, 16)
    ref32 = pipeline.nuc_identity(sh, s, 16)
    assert np.array_equal(a, ref32)
    txt = np.array([float("%.6g" % float(x)) for x in a])          # what ostream << float / stod give the reference
    assert np.array_equal(b, txt)
    assert np.array_equal(b, pipeline.round_6_significant(a))


def test_nloc_batch_matches_numpy():
    """mm_nloc_batch (fEM.h:324-348) against the numpy restatement, with contigs shorter than some reads."""
    import numpy as np
    from metamaps_b200 import capi, pipeline
    from tests.conftest import build_emu
    lib = capi.load(build_emu())
    rng = np.random.default_rng(5)
    n_contigs, T, n_reads = 60, 17, 400
    contig_taxon = rng.integers(0, T, n_contigs).astype(np.int32); contig_taxon[:T] = np.arange(T)
    contig_len = rng.integers(500, 9000, n_contigs).astype(np.int64)
    per = rng.integers(1, 9, n_reads)
    read_off = np.zeros(n_reads + 1, np.int64); read_off[1:] = np.cumsum(per)
    M = int(read_off[-1])
    seq = rng.integers(0, n_contigs, M).astype(np.int32)
    read_len = rng.integers(1000, 8000, n_reads).astype(np.int32)
    tax, nloc = capi.nloc_batch(lib, seq, read_off, read_len, contig_len, contig_taxon, T)
    m_read = np.repeat(np.arange(n_reads), per)
    ref = pipeline._nloc(contig_taxon[seq], seq, m_read, read_len[m_read].astype(np.int64), contig_len, contig_taxon, T)
    assert np.array_equal(tax, contig_taxon[seq])
    assert np.array_equal(nloc, ref)
    # every contig longer than every read: the closed form
    contig_len2 = contig_len + 10000
    tax2, nloc2 = capi.nloc_batch(lib, seq, read_off, read_len, contig_len2, contig_taxon, T)
    ref2 = pipeline._nloc(contig_taxon[seq], seq, m_read, read_len[m_read].astype(np.int64), contig_len2, contig_taxon, T)
    assert np.array_equal(nloc2, ref2)
