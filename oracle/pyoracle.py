"""ctypes front end to the parity checkers.  TEST INFRASTRUCTURE ONLY.

`Oracle`     -> oracle/liboracle.so            (CPU restatement, mm_oracle.cpp)
`RefHarness` -> oracle/_ref/libmm_refharness.so (the unmodified reference headers, ref_harness.cpp)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
Both expose the same method names so that a test can run the same assertions against either.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")


def build(ref: bool = True) -> None:
    """make -C oracle (liboracle.so; the reference-derived files only when /root/reference exists)."""
    subprocess.run(["make", "-C", HERE, "-s", "all" if ref else "liboracle.so"], check=True)


class _Base:
    prefix = ""

    def __init__(self, path: str):
        if not os.path.exists(path):
            build()
        self.lib = C.CDLL(path)
        p = self.prefix
        L = self.lib
        getattr(L, p + "hash").restype = C.c_uint32
        getattr(L, p + "hash").argtypes = [C.c_char_p, C.c_int]
        getattr(L, p + "minimizers").restype = C.c_int64
        getattr(L, p + "minimizers").argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, _u32p, _i32p, _i32p, C.c_int64]
        getattr(L, p + "min_hits_relaxed").restype = C.c_int
        getattr(L, p + "min_hits_relaxed").argtypes = [C.c_int, C.c_int, C.c_float]
        getattr(L, p + "recommended_window").restype = C.c_int
        getattr(L, p + "recommended_window").argtypes = [C.c_double, C.c_int, C.c_int, C.c_float, C.c_int, C.c_uint64]
        getattr(L, p + "identity").restype = None
        getattr(L, p + "identity").argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        getattr(L, p + "index_free").argtypes = [C.c_void_p]
        for nm in ("index_size", "index_unique"):
            getattr(L, p + nm).restype = C.c_int64
            getattr(L, p + nm).argtypes = [C.c_void_p]
        getattr(L, p + "index_freq_threshold").restype = C.c_int
        getattr(L, p + "index_freq_threshold").argtypes = [C.c_void_p]
        getattr(L, p + "index_get").argtypes = [C.c_void_p, _u32p, _i32p, _i32p, _i32p]

    def hash(self, kmer: bytes) -> int:
        return getattr(self.lib, self.prefix + "hash")(kmer, len(kmer))

    def minimizers(self, seq: bytes, k: int, w: int):
        cap = max(16, len(seq))
        h = np.empty(cap, np.uint32); wp = np.empty(cap, np.int32); st = np.empty(cap, np.int32)
        n = getattr(self.lib, self.prefix + "minimizers")(seq, len(seq), k, w, h, wp, st, cap)
        return h[:n].copy(), wp[:n].copy(), st[:n].copy()

    def min_hits_relaxed(self, s: int, k: int, pi: float) -> int:
        return getattr(self.lib, self.prefix + "min_hits_relaxed")(s, k, pi)

    def recommended_window(self, p: float, k: int, pi: float, lenQ: int, lenR: int) -> int:
        return getattr(self.lib, self.prefix + "recommended_window")(p, k, 4, pi, lenQ, lenR)

    def identity(self, shared: int, s: int, k: int):
        a = C.c_float(); b = C.c_float()
        getattr(self.lib, self.prefix + "identity")(shared, s, k, C.byref(a), C.byref(b))
        return a.value, b.value

    # ---- index handle helpers
    def index_size(self, h): return getattr(self.lib, self.prefix + "index_size")(h)
    def index_unique(self, h): return getattr(self.lib, self.prefix + "index_unique")(h)
    def index_freq_threshold(self, h): return getattr(self.lib, self.prefix + "index_freq_threshold")(h)
    def index_free(self, h): getattr(self.lib, self.prefix + "index_free")(h)

    def index_get(self, h):
        n = self.index_size(h)
        hs = np.empty(n, np.uint32); sq = np.empty(n, np.int32); wp = np.empty(n, np.int32); st = np.empty(n, np.int32)
        getattr(self.lib, self.prefix + "index_get")(h, hs, sq, wp, st)
        return hs, sq, wp, st

    def _map_read_call(self, fn, h, seq: bytes, extra, cap: int = 4096):
        info = np.zeros(4, np.int64)
        arrs = [np.zeros(cap, np.int32) for _ in range(7)]
        o1 = np.zeros(cap, np.int64); o2 = np.zeros(cap, np.int64)
        n = fn(h, seq, len(seq), *extra, info, *arrs, o1, o2, cap)
        assert n <= cap
        keys = ["seq", "start", "end", "pos", "shared", "votes", "valid"]
        out = {k: a[:n].copy() for k, a in zip(keys, arrs)}
        out["optStart"] = o1[:n].copy(); out["optEnd"] = o2[:n].copy()
        out["s"] = int(info[0]); out["minimumHits"] = int(info[1]); out["nHits"] = int(info[3])
        return out


class Oracle(_Base):
    prefix = "mmo_"

    def __init__(self):
        super().__init__(os.path.join(HERE, "liboracle.so"))
        L = self.lib
        L.mmo_index_build.restype = C.c_void_p
        L.mmo_index_build.argtypes = [C.c_char_p, _i64p, C.c_int, C.c_int, C.c_int]
        L.mmo_map_read.restype = C.c_int
        L.mmo_map_read.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_float, _i64p] + [_i32p] * 7 + [_i64p, _i64p, C.c_int]
        L.mmo_read_sketch.restype = C.c_int
        L.mmo_read_sketch.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, _u32p, _i32p, _i32p, C.c_int]
        L.mmo_mapq.restype = C.c_int
        L.mmo_mapq.argtypes = [_f64p, _i32p, _i32p, C.c_int, C.c_int, C.c_int, _f64p]
        L.mmo_em.restype = C.c_int
        L.mmo_em.argtypes = [_i32p, _f64p, _f64p, _i64p, C.c_int64, C.c_int, C.c_int, _f64p, _f64p, _i64p, _f64p, C.c_int]
        L.mmo_binom_pmf.restype = C.c_double
        L.mmo_binom_pmf.argtypes = [C.c_int, C.c_int, C.c_double]
        L.mmo_binom_quantile_upper.restype = C.c_double
        L.mmo_binom_quantile_upper.argtypes = [C.c_int, C.c_double, C.c_double]
        L.mmo_binom_sf.restype = C.c_double
        L.mmo_binom_sf.argtypes = [C.c_int, C.c_int, C.c_double]
        L.mmo_estimate_pvalue.restype = C.c_double
        L.mmo_estimate_pvalue.argtypes = [C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_uint64]

    def index_build(self, contigs, k: int, w: int):
        """contigs: list of ASCII bytes."""
        offs = np.zeros(len(contigs) + 1, np.int64)
        offs[1:] = np.cumsum([len(c) for c in contigs])
        return self.lib.mmo_index_build(b"".join(contigs), offs, len(contigs), k, w)

    def map_read(self, h, seq: bytes, pi: float = 80.0, cap: int = 4096):
        return self._map_read_call(self.lib.mmo_map_read, h, seq, (C.c_float(pi),), cap)

    def read_sketch(self, seq: bytes, k: int, w: int):
        cap = max(16, len(seq))
        hs = np.empty(cap, np.uint32); wp = np.empty(cap, np.int32); st = np.empty(cap, np.int32)
        s = self.lib.mmo_read_sketch(seq, len(seq), k, w, hs, wp, st, cap)
        return hs[:s].copy(), wp[:s].copy(), st[:s].copy()

    def mapq(self, identity, shared, sketch, read_len: int, k: int):
        identity = np.ascontiguousarray(identity, np.float64)
        out = np.zeros(len(identity), np.float64)
        rc = self.lib.mmo_mapq(identity, np.ascontiguousarray(shared, np.int32), np.ascontiguousarray(sketch, np.int32),
                               len(identity), read_len, k, out)
        return rc, out

    def em(self, taxon, mapq, nloc, read_off, T: int, max_iter: int = 0):
        taxon = np.ascontiguousarray(taxon, np.int32); mapq = np.ascontiguousarray(mapq, np.float64)
        nloc = np.ascontiguousarray(nloc, np.float64); read_off = np.ascontiguousarray(read_off, np.int64)
        nr = len(read_off) - 1
        f = np.zeros(T); post = np.zeros(len(taxon)); best = np.zeros(nr, np.int64); ll = np.zeros(4096)
        it = self.lib.mmo_em(taxon, mapq, nloc, read_off, nr, T, max_iter, f, post, best, ll, len(ll))
        return {"f": f, "posterior": post, "best": best, "ll": ll[:min(it, len(ll))].copy(), "iters": it}

    def binom_pmf(self, k, n, p): return self.lib.mmo_binom_pmf(k, n, p)
    def binom_quantile_upper(self, n, p, q): return self.lib.mmo_binom_quantile_upper(n, p, q)
    def binom_sf(self, k, n, p): return self.lib.mmo_binom_sf(k, n, p)
    def estimate_pvalue(self, s, k, pi, lenQ, lenR): return self.lib.mmo_estimate_pvalue(s, k, 4, pi, lenQ, lenR)


REF_SO = os.path.join(HERE, "_ref", "libmm_refharness.so")
REF_BIN = os.path.join(HERE, "_ref", "metamaps")


def ref_available() -> bool:
    return os.path.exists(REF_SO) and os.path.exists(REF_BIN)


class RefHarness(_Base):
    prefix = "mmr_"

    def __init__(self):
        super().__init__(REF_SO)
        L = self.lib
        L.mmr_index_build.restype = C.c_void_p
        L.mmr_index_build.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_float]
        L.mmr_map_read.restype = C.c_int
        L.mmr_map_read.argtypes = [C.c_void_p, C.c_char_p, C.c_int, _i64p] + [_i32p] * 7 + [_i64p, _i64p, C.c_int]
        L.mmr_read_sketch.restype = C.c_int
        L.mmr_read_sketch.argtypes = [C.c_void_p, C.c_char_p, C.c_int, _u32p, _i32p, _i32p, C.c_int]
        L.mmr_likelihood.restype = C.c_double
        L.mmr_likelihood.argtypes = [C.c_int, C.c_int, C.c_double, C.c_int, C.c_int]

    def index_build_fasta(self, fasta: str, k: int, w: int, min_read_len: int = 1000, pi: float = 80.0):
        return self.lib.mmr_index_build(fasta.encode(), k, w, min_read_len, pi)

    def map_read(self, h, seq: bytes, pi: float = 80.0, cap: int = 4096):
        return self._map_read_call(self.lib.mmr_map_read, h, seq, (), cap)

    def read_sketch(self, h, seq: bytes):
        cap = max(16, len(seq))
        hs = np.empty(cap, np.uint32); wp = np.empty(cap, np.int32); st = np.empty(cap, np.int32)
        s = self.lib.mmr_read_sketch(h, seq, len(seq), hs, wp, st, cap)
        return hs[:s].copy(), wp[:s].copy(), st[:s].copy()

    def likelihood(self, k, n_kmers, identity, sketch, inter):
        return self.lib.mmr_likelihood(k, n_kmers, identity, sketch, inter)
