"""ctypes binding of the C ABI in include/metamaps_b200.h.

`load()` binds metamaps_b200/libmetamaps_b200.so (the nvcc sm_100a build).  There is no Python or CPU
implementation behind it: if the shared library is missing, or no CUDA device is usable, the calls fail.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmetamaps_b200.so")

_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")


class MapParams(C.Structure):
    _fields_ = [("perc_identity", C.c_float), ("min_read_len", C.c_int32), ("report_all", C.c_int32), ("reserved", C.c_int32)]


class MapSummary(C.Structure):
    _fields_ = [("n_reads", C.c_int64), ("n_too_short", C.c_int64), ("n_candidates", C.c_int64), ("n_mappings", C.c_int64),
                ("n_reads_mapped", C.c_int64), ("total_bases_mapped_reads", C.c_int64)]


class ClassifySummary(C.Structure):
    _fields_ = [("n_mappings", C.c_int64), ("n_reads_mapped", C.c_int64), ("em_iters", C.c_int32), ("n_identity_fixups", C.c_int32),
                ("em_ms", C.c_double), ("classify_ms", C.c_double)]


# every symbol include/metamaps_b200.h declares: (restype, argtypes); None = opaque/void pointers
SYMBOLS = {
    "mm_last_error": (C.c_char_p, []),
    "mm_version": (C.c_char_p, []),
    "mm_ctx_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "mm_ctx_destroy": (None, [C.c_void_p]),
    "mm_ctx_last_timing": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "mm_ctx_last_map_stats": (C.c_int, [C.c_void_p, _f64p, _i64p]),
    "mm_sketch_batch": (C.c_int, [C.c_void_p, C.c_char_p, _i64p, C.c_int32, C.c_int, C.c_int, C.POINTER(C.c_int64)]),
    "mm_sketch_fetch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mm_index_create": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "mm_index_add": (C.c_int, [C.c_void_p, C.c_char_p, _i64p, C.c_int32]),
    "mm_index_add_dev": (C.c_int, [C.c_void_p, C.c_void_p, _i64p, C.c_int32]),
    "mm_index_set_shard": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32]),
    "mm_index_sync_threshold": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int64)]),
    "mm_index_set_freq_carry": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]),
    "mm_index_get_freq_hist": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "mm_comm_set_rank": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "mm_ctx_mem_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "mm_index_save": (C.c_int, [C.c_void_p, C.c_char_p]),
    "mm_index_load": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p)]),
    "mm_index_params": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_void_p]),
    "mm_index_finalize": (C.c_int, [C.c_void_p]),
    "mm_index_max_batch_reads": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
    "mm_index_stats": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int64)]),
    "mm_index_fetch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mm_index_lookup": (C.c_int, [C.c_void_p, _u32p, C.c_int64, _i32p]),
    "mm_index_destroy": (None, [C.c_void_p]),
    "mm_map_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_char_p, _i64p, C.c_int32, C.POINTER(MapParams), C.POINTER(MapSummary)]),
    "mm_map_batch_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, _i64p, C.c_int32, C.POINTER(MapParams), C.POINTER(MapSummary)]),
    "mm_map_batch_sharded_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, _i64p, C.c_int32, C.POINTER(MapParams), C.POINTER(MapSummary), C.POINTER(C.c_int64)]),
    "mm_stage_reads_async": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, _i64p, C.c_int32]),
    "mm_map_batch_staged": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(MapParams), C.POINTER(MapSummary)]),
    "mm_map_fetch_reads": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mm_map_fetch_candidates": (C.c_int, [C.c_void_p] + [C.c_void_p] * 10),
    "mm_map_fetch_sketch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]),
    "mm_map_fetch_mappings": (C.c_int, [C.c_void_p] + [C.c_void_p] * 8 + [C.c_int64, C.POINTER(C.c_int64)]),
    "mm_group_sorted": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.POINTER(C.c_int64)]),
    "mm_stat_identity_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]),
    "mm_nloc_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "mm_stat_min_hits_relaxed": (C.c_int, [C.c_int, C.c_int, C.c_float]),
    "mm_stat_recommended_window": (C.c_int, [C.c_double, C.c_int, C.c_int, C.c_float, C.c_int, C.c_uint64]),
    "mm_stat_estimate_pvalue": (C.c_double, [C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_uint64]),
    "mm_stat_identity": (None, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "mm_mapq_batch": (C.c_int, [C.c_void_p, _f64p, _i32p, _i32p, _i32p, _i64p, C.c_int64, C.c_int, _f64p, _i32p]),
    "mm_em_run": (C.c_int, [C.c_void_p, _i32p, _f64p, _f64p, _i64p, C.c_int64, C.c_int32, C.c_int32, _f64p, _f64p, _i64p, _f64p, C.c_int32, C.POINTER(C.c_int32)]),
    "mm_classify_setup": (C.c_int, [C.c_void_p, _i64p, _i32p, C.c_int32, C.c_int32]),
    "mm_classify_begin": (C.c_int, [C.c_void_p]),
    "mm_classify_add_mappings": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_int64)]),
    "mm_classify_next_batch": (C.c_int, [C.c_void_p]),
    "mm_classify_exchange": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_int64)]),
    "mm_classify_run": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(ClassifySummary)]),
    "mm_classify_fetch": (C.c_int, [C.c_void_p] + [C.c_void_p] * 12 + [C.c_int64] + [C.c_void_p] * 6 + [C.c_int32]),
    "mm_comm_unique_id": (C.c_int, [C.c_void_p]),
    "mm_comm_init": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "mm_comm_destroy": (C.c_int, [C.c_void_p]),
    "mm_comm_set_allreduce": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
}
ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.POINTER(C.c_double), C.c_int64, C.c_void_p)


class MMError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"metamaps_b200 error {code}: {msg}")
        self.code = code


def load(path: str | None = None) -> C.CDLL:
    """dlopen the C-ABI library and declare every prototype.  Raises if the library is absent."""
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise FileNotFoundError(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  There is no CPU fallback.")
    lib = C.CDLL(path)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _ascii_batch(seqs):
    """list of bytes -> (joined bytes, int64 offsets)."""
    offs = np.zeros(len(seqs) + 1, np.int64)
    if len(seqs):
        offs[1:] = np.cumsum([len(s) for s in seqs])
    return b"".join(seqs), offs


class Context:
    """One per GPU (mm_ctx)."""

    def __init__(self, device: int = 0, lib: C.CDLL | None = None):
        self.lib = lib or load()
        h = C.c_void_p()
        self._check(self.lib.mm_ctx_create(device, C.byref(h)))
        self.h = h
        self.device = device

    def _check(self, rc: int):
        if rc != 0:
            raise MMError(rc, (self.lib.mm_last_error() or b"").decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.mm_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def last_timing(self):
        ms = C.c_double(); n = C.c_int64()
        self.lib.mm_ctx_last_timing(self.h, C.byref(ms), C.byref(n))
        return ms.value, n.value

    def last_map_stats(self):
        ms = np.zeros(16); ct = np.zeros(16, np.int64)
        self.lib.mm_ctx_last_map_stats(self.h, ms, ct)
        return {"sketch_ms": ms[0], "read_sketch_ms": ms[1], "l1_probe_ms": ms[2], "l1_sort_ms": ms[3], "l1_candidates_ms": ms[4],
                "l2_setup_ms": ms[5], "l2_classify_ms": ms[6], "l2_sweep_ms": ms[7], "l2_strand_ms": ms[8], "accept_ms": ms[9],
                "sketch_elems": int(ct[0]), "hits": int(ct[1]), "candidates": int(ct[2]), "span_elems": int(ct[3]),
                "mappings": int(ct[4]), "read_minimizers": int(ct[5]), "bases": int(ct[6]), "exceptions": int(ct[7]),
                "ambiguous_reads": int(ct[8]), "smem_swept": int(ct[9]), "hits_kept": int(ct[10]), "sweep_items": int(ct[11]),
                "k1_kernel_ms": ms[10], "sweep_kernel_ms": ms[11], "l2_prune_ms": ms[12], "l1_kernel_ms": ms[13],
                "window_starts_swept": int(ct[12]), "window_starts": int(ct[13])}

    # K1
    def sketch(self, seqs, k: int, w: int):
        """CommonFunc::addMinimizers over a list of ASCII sequences -> (offsets, hash, wpos, strand)."""
        data, offs = _ascii_batch(seqs)
        n = C.c_int64()
        self._check(self.lib.mm_sketch_batch(self.h, data, offs, len(seqs), k, w, C.byref(n)))
        counts = np.zeros(len(seqs) + 1, np.int64)
        hs = np.zeros(n.value, np.uint32); wp = np.zeros(n.value, np.int32); st = np.zeros(n.value, np.int32)
        self._check(self.lib.mm_sketch_fetch(self.h, _ptr(counts), _ptr(hs), _ptr(wp), _ptr(st)))
        return counts, hs, wp, st

    # K6
    def mapq(self, identity, shared, sketch, read_len, read_off, k: int):
        identity = np.ascontiguousarray(identity, np.float64)
        out = _out("mapq", len(identity), np.float64); status = _out("mapq_status", len(read_off) - 1, np.int32)
        self._check(self.lib.mm_mapq_batch(self.h, identity, np.ascontiguousarray(shared, np.int32),
                                           np.ascontiguousarray(sketch, np.int32), np.ascontiguousarray(read_len, np.int32),
                                           np.ascontiguousarray(read_off, np.int64), len(read_off) - 1, k, out, status))
        return out, status

    # K7/K8
    def em(self, taxon, mapq, nloc, read_off, T: int, max_iter: int = 0):
        taxon = np.ascontiguousarray(taxon, np.int32); mapq = np.ascontiguousarray(mapq, np.float64)
        nloc = np.ascontiguousarray(nloc, np.float64); read_off = np.ascontiguousarray(read_off, np.int64)
        nr = len(read_off) - 1
        f = _out("em_f", T, np.float64); post = _out("em_post", len(taxon), np.float64); best = _out("em_best", max(nr, 1), np.int64); ll = np.zeros(4096)
        it = C.c_int32()
        self._check(self.lib.mm_em_run(self.h, taxon, mapq, nloc, read_off, nr, T, max_iter, f, post, best, ll, len(ll), C.byref(it)))
        return {"f": f, "posterior": post, "best": best[:nr], "ll": ll[:min(it.value, len(ll))].copy(), "iters": it.value}

    # classify stage on the device (mm_classify_*)
    def classify_setup(self, contig_len, contig_taxon, n_taxa: int):
        """Contig lengths / taxa (global contig ids) of the reference; cached on the arrays' identity."""
        key = (id(contig_len), id(contig_taxon), int(n_taxa))
        if getattr(self, "_taxo_key", None) == key:
            return
        cl = np.ascontiguousarray(contig_len, np.int64); ct = np.ascontiguousarray(contig_taxon, np.int32)
        self._check(self.lib.mm_classify_setup(self.h, cl, ct, len(cl), int(n_taxa)))
        self._taxo_key = key; self._taxo_keep = (contig_len, contig_taxon); self._n_taxa = int(n_taxa)

    def classify_begin(self):
        self._check(self.lib.mm_classify_begin(self.h))

    def classify_add(self, first_contig_id: int = 0) -> int:
        """Append the accepted mappings of the last map call (contig ids + first_contig_id); returns the table size."""
        n = C.c_int64()
        self._check(self.lib.mm_classify_add_mappings(self.h, int(first_contig_id), C.byref(n)))
        return n.value

    def classify_next_batch(self):
        """The next map call's reads are numbered after the ones already in the table (several batches, one EM)."""
        self._check(self.lib.mm_classify_next_batch(self.h))

    def classify_exchange(self, read_lo: int, read_hi: int) -> int:
        """Collective (contig-sharded ranks): all-gather + merge the tables, keep the reads [read_lo, read_hi)."""
        n = C.c_int64()
        self._check(self.lib.mm_classify_exchange(self.h, int(read_lo), int(read_hi), C.byref(n)))
        return n.value

    def classify_run(self, em_max_iter: int = 0) -> dict:
        s = ClassifySummary()
        self._check(self.lib.mm_classify_run(self.h, int(em_max_iter), C.byref(s)))
        return {f[0]: getattr(s, f[0]) for f in ClassifySummary._fields_}

    def classify_fetch(self, summary: dict, what=("read", "seq", "pos", "shared", "sketch", "strand", "identity", "identity_parsed", "mapq", "taxon",
                                                    "nloc", "posterior", "mapped_reads", "read_off", "best", "mapq_status", "f", "ll")) -> dict:
        """D2H of the finished arrays (one call, one synchronisation)."""
        M = int(summary["n_mappings"]); G = int(summary["n_reads_mapped"]); T = self._n_taxa
        per_m = (("read", np.int32), ("seq", np.int32), ("pos", np.int32), ("shared", np.int32), ("sketch", np.int32), ("strand", np.int32),
                 ("identity", np.float32), ("identity_parsed", np.float64), ("mapq", np.float64), ("taxon", np.int32), ("nloc", np.float64),
                 ("posterior", np.float64))
        out = {}; args = []
        for name, dt in per_m:
            a = _out("c_" + name, M, dt) if name in what else None
            out[name] = a; args.append(_ptr(a))
        mr = _out("c_mapped", G, np.int32) if "mapped_reads" in what else None
        ro = _out("c_read_off", G + 1, np.int64) if "read_off" in what else None
        be = _out("c_best", G, np.int64) if "best" in what else None
        stt = _out("c_status", G, np.int32) if "mapq_status" in what else None
        f = _out("c_f", T, np.float64) if "f" in what else None
        ll = np.zeros(4096) if "ll" in what else None
        self._check(self.lib.mm_classify_fetch(self.h, *args, M, _ptr(mr), _ptr(ro), _ptr(be), _ptr(stt), _ptr(f), _ptr(ll), 4096 if ll is not None else 0))
        out.update({"mapped_reads": mr, "read_off": ro, "best": be, "mapq_status": stt, "f": f,
                    "ll": None if ll is None else ll[:min(int(summary["em_iters"]), 4096)].copy()})
        out["d2h_bytes"] = int(sum(v.nbytes for v in out.values() if isinstance(v, np.ndarray)))
        return {k_: v for k_, v in out.items() if v is not None}

    def comm_init(self, n_ranks: int, rank: int, uid: bytes):
        buf = C.create_string_buffer(uid, 128)
        self._check(self.lib.mm_comm_init(self.h, n_ranks, rank, buf))
        self.n_ranks_hint = n_ranks

    def set_allreduce(self, fn):
        """fn(numpy float64 array) -> None must sum-all-reduce the array in place across ranks (host transport)."""
        if fn is None:
            self._ar = None
            self._check(self.lib.mm_comm_set_allreduce(self.h, None, None))
            return

        def tramp(ptr, n, user):
            try:
                fn(np.ctypeslib.as_array(ptr, shape=(n,)))
                return 0
            except Exception:
                return 1
        self._ar = ALLREDUCE_FN(tramp)        # keep the trampoline alive
        self._check(self.lib.mm_comm_set_allreduce(self.h, C.cast(self._ar, C.c_void_p), None))

    def stage_reads(self, slot: int, host_ptr: int, offsets):
        """Start the H2D copy of a batch (one host buffer, ideally pinned) into staging slot 0/1; returns at once."""
        offs = np.ascontiguousarray(offsets, np.int64)
        self._check(self.lib.mm_stage_reads_async(self.h, slot, C.c_void_p(host_ptr), offs, len(offs) - 1))

    n_ranks_hint = 1       # > 1 once a communicator / host transport is attached: collective calls must not be skipped

    def set_rank(self, n_ranks: int, rank: int):
        self._check(self.lib.mm_comm_set_rank(self.h, n_ranks, rank))
        self.n_ranks_hint = n_ranks

    def comm_unique_id(self) -> bytes:
        buf = C.create_string_buffer(128)
        self._check(self.lib.mm_comm_unique_id(buf))
        return buf.raw


class Index:
    """skch::Sketch on the GPU (mm_index)."""

    def __init__(self, ctx: Context, k: int, w: int):
        self.ctx = ctx; self.lib = ctx.lib; self.k = k; self.w = w
        h = C.c_void_p()
        ctx._check(self.lib.mm_index_create(ctx.h, k, w, C.byref(h)))
        self.h = h

    def add(self, contigs):
        data, offs = _ascii_batch(contigs)
        self.ctx._check(self.lib.mm_index_add(self.h, data, offs, len(contigs)))

    def add_dev(self, dev_ptr: int, offsets):
        offs = np.ascontiguousarray(offsets, np.int64)
        self.ctx._check(self.lib.mm_index_add_dev(self.h, C.c_void_p(dev_ptr), offs, len(offs) - 1))

    def finalize(self):
        self.ctx._check(self.lib.mm_index_finalize(self.h))

    def save(self, path: str):
        """GPU-native dump of the finalized index (mm_index_save)."""
        self.ctx._check(self.lib.mm_index_save(self.h, path.encode()))

    @classmethod
    def load(cls, ctx: "Context", path: str) -> "Index":
        """mm_index_load: the arrays go straight back to the device."""
        self = cls.__new__(cls)
        self.ctx = ctx; self.lib = ctx.lib
        h = C.c_void_p()
        ctx._check(ctx.lib.mm_index_load(ctx.h, path.encode(), C.byref(h)))
        self.h = h
        k = C.c_int32(); w = C.c_int32()
        ctx._check(ctx.lib.mm_index_params(h, C.byref(k), C.byref(w), None))
        self.k, self.w = k.value, w.value
        return self

    def set_shard(self, first_contig_id: int, keep_counts: bool = True):
        """This index holds the contigs [first_contig_id, ...) of a larger reference (call before finalize)."""
        self.first_contig = int(first_contig_id)
        self.ctx._check(self.lib.mm_index_set_shard(self.h, int(first_contig_id), 1 if keep_counts else 0))

    def set_freq_carry(self, hist, prev_threshold: int):
        """hist: list of (occurrence count, number of hashes) carried from the previous reference chunks (see the header)."""
        v = np.array([h[0] for h in hist], np.int32); c = np.array([h[1] for h in hist], np.int64)
        self.ctx._check(self.lib.mm_index_set_freq_carry(self.h, _ptr(v), _ptr(c), len(v), int(prev_threshold)))

    def freq_hist(self):
        """(cumulative histogram after this chunk as a list of (count, hashes), this chunk's threshold)."""
        n = C.c_int32(); t = C.c_int32()
        self.ctx._check(self.lib.mm_index_get_freq_hist(self.h, None, None, 0, C.byref(n), C.byref(t)))
        v = np.zeros(n.value, np.int32); c = np.zeros(n.value, np.int64)
        self.ctx._check(self.lib.mm_index_get_freq_hist(self.h, _ptr(v), _ptr(c), n.value, C.byref(n), C.byref(t)))
        return [(int(a), int(b)) for a, b in zip(v, c)], t.value

    def sync_threshold(self):
        """Collective: occurrence threshold of the whole (sharded) reference; flags the over-frequent hashes locally."""
        t = C.c_int32(); u = C.c_int64()
        self.ctx._check(self.lib.mm_index_sync_threshold(self.h, C.byref(t), C.byref(u)))
        return t.value, u.value

    def max_batch_reads(self) -> int:
        n = C.c_int64()
        self.ctx._check(self.lib.mm_index_max_batch_reads(self.h, C.byref(n)))
        return n.value

    def stats(self):
        a = C.c_int64(); b = C.c_int64(); c = C.c_int32(); d = C.c_int32(); e = C.c_int64()
        self.ctx._check(self.lib.mm_index_stats(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(d), C.byref(e)))
        return {"n_minimizers": a.value, "n_unique": b.value, "freq_threshold": c.value, "n_contigs": d.value, "device_bytes": e.value}

    def fetch(self):
        n = self.stats()["n_minimizers"]
        hs = np.zeros(n, np.uint32); sq = np.zeros(n, np.int32); wp = np.zeros(n, np.int32); st = np.zeros(n, np.int32)
        self.ctx._check(self.lib.mm_index_fetch(self.h, _ptr(hs), _ptr(sq), _ptr(wp), _ptr(st)))
        return hs, sq, wp, st

    def lookup(self, hashes):
        hashes = np.ascontiguousarray(hashes, np.uint32)
        out = np.zeros(len(hashes), np.int32)
        self.ctx._check(self.lib.mm_index_lookup(self.h, hashes, len(hashes), out))
        return out

    def close(self):
        if getattr(self, "h", None):
            self.lib.mm_index_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def fetch_map_results(ctx: Context, out: dict) -> dict:
    """D2H of the per-read and per-candidate results of the last map call (mm_map_fetch_reads / _candidates)."""
    n, nc = out["_n"], out["_nc"]
    sk = np.zeros(n, np.int32); mh = np.zeros(n, np.int32); co = np.zeros(n + 1, np.int64)
    ctx._check(ctx.lib.mm_map_fetch_reads(ctx.h, _ptr(sk), _ptr(mh), _ptr(co)))
    names = ["seq", "start", "end", "pos", "shared", "votes", "accepted", "valid"]
    arrs = [np.zeros(nc, np.int32) for _ in names]
    o1 = np.zeros(nc, np.int64); o2 = np.zeros(nc, np.int64)
    ctx._check(ctx.lib.mm_map_fetch_candidates(ctx.h, *[_ptr(a) for a in arrs], _ptr(o1), _ptr(o2)))
    out.update({"s": sk, "minimumHits": mh, "cand_off": co, "optStart": o1, "optEnd": o2})
    out.update(dict(zip(names, arrs)))
    out["d2h_bytes"] = int(sk.nbytes + mh.nbytes + co.nbytes + sum(a.nbytes for a in arrs) + o1.nbytes + o2.nbytes)
    return out


class _PinnedPool:
    """Grow-only pinned host buffers for result arrays, one per (thread, name): a streaming host reuses its result buffers
    from batch to batch, and device<->host copies to pinned memory run at full PCIe speed instead of being staged."""

    def __init__(self):
        import torch
        self.torch = torch
        self.bufs = {}

    def get(self, name: str, n: int, dtype):
        import threading
        key = (threading.get_ident(), name)
        need = max(int(n) * np.dtype(dtype).itemsize, 8)
        buf = self.bufs.get(key)
        if buf is None or buf.numel() < need:
            buf = self.torch.empty(max(need + need // 4, 1 << 16), dtype=self.torch.uint8, pin_memory=True)
            self.bufs[key] = buf
        return buf.numpy()[:int(n) * np.dtype(dtype).itemsize].view(dtype)


_POOL = None


def use_pinned_results(flag: bool = True):
    """Result arrays of fetch_mappings / mapq / nloc_batch / em come from reused pinned buffers (valid until the same
    thread's next call of the same function) instead of fresh numpy arrays."""
    global _POOL
    _POOL = _PinnedPool() if flag else None


def _out(name: str, n: int, dtype):
    return _POOL.get(name, n, dtype) if _POOL is not None else np.zeros(int(n), dtype)


def fetch_mappings(ctx: Context, n_mappings: int) -> dict:
    """The accepted mappings of the last map call, compacted on the device (mm_map_fetch_mappings)."""
    n = int(n_mappings)
    out = {k_: _out("m_" + k_, n, np.int32) for k_ in ("read", "seq", "pos", "shared", "sketch", "strand")}
    out["identity"] = _out("m_identity", n, np.float32); out["identity_parsed"] = _out("m_identity_parsed", n, np.float64)
    got = C.c_int64()
    ctx._check(ctx.lib.mm_map_fetch_mappings(ctx.h, *[_ptr(out[k_]) for k_ in ("read", "seq", "pos", "shared", "sketch", "strand", "identity",
                                                                                "identity_parsed")], n, C.byref(got)))
    assert got.value == n, (got.value, n)
    out["d2h_bytes"] = int(6 * 4 * n)
    return out


def group_sorted(lib, values):
    """(distinct values, offsets) of the runs of a non-decreasing int32 array (mm_group_sorted)."""
    values = np.ascontiguousarray(values, np.int32)
    gv = _out("grp_value", len(values), np.int32); go = _out("grp_off", len(values) + 1, np.int64)
    ng = C.c_int64()
    rc = lib.mm_group_sorted(_ptr(values), len(values), _ptr(gv), _ptr(go), C.byref(ng))
    if rc != 0:
        raise MMError(rc, (lib.mm_last_error() or b"").decode())
    return gv[:ng.value], go[:ng.value + 1]


def identity_batch(lib, shared, sketch, k: int):
    shared = np.ascontiguousarray(shared, np.int32); sketch = np.ascontiguousarray(sketch, np.int32)
    a = np.zeros(len(shared), np.float32); b = np.zeros(len(shared), np.float64)
    rc = lib.mm_stat_identity_batch(_ptr(shared), _ptr(sketch), len(shared), k, _ptr(a), _ptr(b))
    if rc != 0:
        raise MMError(rc, (lib.mm_last_error() or b"").decode())
    return a, b


def nloc_batch(lib, seq, read_off, read_len, contig_len, contig_taxon, n_taxa: int):
    seq = np.ascontiguousarray(seq, np.int32); read_off = np.ascontiguousarray(read_off, np.int64)
    read_len = np.ascontiguousarray(read_len, np.int32); contig_len = np.ascontiguousarray(contig_len, np.int64)
    contig_taxon = np.ascontiguousarray(contig_taxon, np.int32)
    tax = _out("nloc_tax", len(seq), np.int32); nloc = _out("nloc", len(seq), np.float64)
    rc = lib.mm_nloc_batch(_ptr(seq), _ptr(read_off), _ptr(read_len), len(read_off) - 1, _ptr(contig_len), _ptr(contig_taxon), len(contig_len),
                           int(n_taxa), _ptr(tax), _ptr(nloc))
    if rc != 0:
        raise MMError(rc, (lib.mm_last_error() or b"").decode())
    return tax, nloc


def map_reads_sharded(ctx: Context, index: Index, dev_ptr: int, offsets, perc_identity: float = 80.0, min_read_len: int = 1000):
    """Collective over contig-sharded ranks: `dev_ptr` / `offsets` are THIS rank's block of the batch (device ASCII); every rank ends
    with the map results of the whole batch against its shard (mm_map_batch_sharded_dev).  Returns like map_reads(fetch=False) plus
    "first_read" = index of this rank's first read in the batch."""
    p = MapParams(perc_identity, min_read_len, 1, 0); s = MapSummary(); first = C.c_int64()
    offs = np.ascontiguousarray(offsets, np.int64)
    ctx._check(ctx.lib.mm_map_batch_sharded_dev(ctx.h, index.h, C.c_void_p(dev_ptr), offs, len(offs) - 1, C.byref(p), C.byref(s), C.byref(first)))
    gpu_ms, launches = ctx.last_timing()
    return {"summary": {f[0]: getattr(s, f[0]) for f in MapSummary._fields_}, "gpu_ms": gpu_ms, "launches": launches, "stats": ctx.last_map_stats(),
            "_n": int(s.n_reads), "_nc": int(s.n_candidates), "first_read": first.value}


def map_reads(ctx: Context, index: Index, reads=None, perc_identity: float = 80.0, min_read_len: int = 1000,
              dev_ptr: int | None = None, host_ptr: int | None = None, offsets=None, fetch: bool = True,
              fetch_sketch: bool = False, staged_slot: int | None = None):
    """skch::Map over a batch.  reads: list of ASCII bytes (host); or host_ptr+offsets (one host buffer, e.g.
    pinned); or dev_ptr+offsets (device-resident ASCII)."""
    p = MapParams(perc_identity, min_read_len, 1, 0)
    s = MapSummary()
    if staged_slot is not None:
        offs = np.ascontiguousarray(offsets, np.int64)
        ctx._check(ctx.lib.mm_map_batch_staged(ctx.h, index.h, staged_slot, C.byref(p), C.byref(s)))
    elif host_ptr is not None:
        offs = np.ascontiguousarray(offsets, np.int64)
        ctx._check(ctx.lib.mm_map_batch(ctx.h, index.h, C.cast(host_ptr, C.c_char_p), offs, len(offs) - 1, C.byref(p), C.byref(s)))
    elif dev_ptr is None:
        data, offs = _ascii_batch(reads)
        ctx._check(ctx.lib.mm_map_batch(ctx.h, index.h, data, offs, len(reads), C.byref(p), C.byref(s)))
    else:
        offs = np.ascontiguousarray(offsets, np.int64)
        ctx._check(ctx.lib.mm_map_batch_dev(ctx.h, index.h, C.c_void_p(dev_ptr), offs, len(offs) - 1, C.byref(p), C.byref(s)))
    n = len(offs) - 1
    gpu_ms, launches = ctx.last_timing()          # of the map call itself (the fetches below are separate calls)
    out = {"summary": {f[0]: getattr(s, f[0]) for f in MapSummary._fields_}, "gpu_ms": gpu_ms, "launches": launches,
           "stats": ctx.last_map_stats()}
    out["_n"] = n; out["_nc"] = int(s.n_candidates)
    if not fetch:
        return out
    out = fetch_map_results(ctx, out)
    if fetch_sketch:
        qo = np.zeros(n + 1, np.int64)
        ctx._check(ctx.lib.mm_map_fetch_sketch(ctx.h, _ptr(qo), None, None, 0))
        nq = int(qo[-1])
        qh = np.zeros(nq, np.uint32); qs = np.zeros(nq, np.int32)
        ctx._check(ctx.lib.mm_map_fetch_sketch(ctx.h, _ptr(qo), _ptr(qh), _ptr(qs), nq))
        out.update({"q_off": qo, "q_hash": qh, "q_strand": qs})
    return out
