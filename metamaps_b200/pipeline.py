"""Host-side mirror of the reference's mapDirectly -> classify data flow on arrays (no files).

`map_and_classify` strings the C-ABI calls together the way `metamaps mapDirectly` + `metamaps classify` do:
  skch::Map (K1,K3,K4,K5)  ->  identity + filter (computeMap.hpp:405-415)  ->  addMappingQualities (K6)
  ->  getMappingLocations nloc (fEM.h:324-348)  ->  doEM (K7/K8).
The float arithmetic that fixes the printed identity (and therefore what `classify` re-parses from the text
file, mapWrap.h:229, fEM.h:297) is done here on the host, vectorised, exactly as the reference does it.
"""
from __future__ import annotations

import numpy as np

from . import capi


def nuc_identity(shared: np.ndarray, s: np.ndarray, k: int) -> np.ndarray:
    """float nucIdentity = 100 * (1 - j2md(1.0*shared/s, k))   (map_stats.hpp:44-54, computeMap.hpp:403-408)."""
    j = (shared.astype(np.float64) / s.astype(np.float64)).astype(np.float32)
    jd = j.astype(np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        # `2.0 * j/(1+j)`: (1+j) is evaluated in float (int + float), the rest in double
        md = ((-1.0 / k) * np.log(2.0 * jd / (np.float32(1) + j).astype(np.float64))).astype(np.float32)
    md = np.where(j == 0, np.float32(1.0), np.where(j == 1, np.float32(0.0), md)).astype(np.float32)
    return (np.float32(100) * (np.float32(1) - md)).astype(np.float32)


def round_6_significant(x32: np.ndarray) -> np.ndarray:
    """The double obtained by printing a float with the default ostream precision (6 significant digits) and
    parsing it back (what mapWrap.h:229 / fEM.h:297 see).  x*10^d has at most 24+14 significant bits, so the
    scaling is exact in double and rint()/division give the correctly rounded decimal."""
    x = x32.astype(np.float64)
    if x.size and x.min() >= 10.0 and x.max() < 99.99995:      # the usual case: identities in [10,100) -> 4 decimals
        return np.rint(x * 1e4) / 1e4
    out = np.zeros_like(x)
    nz = x != 0
    e = np.floor(np.log10(np.abs(x[nz])))
    e = np.where(np.abs(x[nz]) >= 10.0 ** (e + 1), e + 1, e)
    e = np.where(np.abs(x[nz]) < 10.0 ** e, e - 1, e)
    d = (5 - e).astype(np.int64)
    scale = 10.0 ** np.abs(d)
    scaled = np.where(d >= 0, x[nz] * scale, x[nz] / scale)
    r = np.rint(scaled)
    out[nz] = np.where(d >= 0, r / scale, r * scale)
    return out


def classify_mappings(ctx: capi.Context, m: dict, n_reads: int, read_len: np.ndarray, *, k: int, contig_len: np.ndarray, contig_taxon: np.ndarray,
                      n_taxa: int, em_max_iter: int = 0, wall: dict | None = None) -> dict:
    """addMappingQualities (K6) -> nLoc (host) -> doEM (K7/K8) over the accepted mappings `m` (capi.fetch_mappings layout,
    global contig ids, read order)."""
    import time
    t2 = time.perf_counter()
    m_read = m["read"]
    # reads with >= 1 mapping, in read order (the mappings file has no lines for the others)
    mapped, read_off = capi.group_sorted(ctx.lib, m_read)
    out = {"read": m_read, "seq": m["seq"], "pos": m["pos"], "shared": m["shared"], "sketch": m["sketch"],
           "strand": m["strand"], "identity": m["identity"], "mapped_reads": mapped, "read_off": read_off}
    if len(m_read) == 0 and ctx.n_ranks_hint <= 1:
        out.update({"mapq": np.zeros(0), "em": None, "gpu_ms": 0.0, "launches": 0, "d2h_bytes": 0})
        return out
    t3 = time.perf_counter()
    rl_mapped = np.ascontiguousarray(read_len[mapped], np.int32)
    ident = np.divide(m["identity_parsed"], 100.0, out=capi._out("ident_frac", len(m_read), np.float64))    # column 10 / 100 (mapWrap.h:229)
    mapq, status = ctx.mapq(ident, m["shared"], m["sketch"], rl_mapped, read_off, k)
    t4 = time.perf_counter()
    launches = ctx.last_timing()[1]; gpu_ms = ctx.last_timing()[0]
    out["mapq"] = mapq; out["mapq_status"] = status
    # fEM.h:324-348: possible mapping locations of the mapping's taxon for this read length
    tax, nloc = capi.nloc_batch(ctx.lib, m["seq"], read_off, rl_mapped, contig_len, contig_taxon, n_taxa)
    t5 = time.perf_counter()
    em = ctx.em(tax, mapq, nloc, read_off, n_taxa, em_max_iter)
    t6 = time.perf_counter()
    if wall is not None:
        wall.update({"read_offsets": (t3 - t2) * 1e3, "mapq_call": (t4 - t3) * 1e3, "nloc": (t5 - t4) * 1e3, "em_call": (t6 - t5) * 1e3})
    launches += ctx.last_timing()[1]; gpu_ms += ctx.last_timing()[0]
    d2h = mapq.nbytes + status.nbytes + em["f"].nbytes + em["posterior"].nbytes + em["best"].nbytes
    out.update({"taxon": tax, "nloc": nloc, "em": em, "gpu_ms": gpu_ms, "launches": launches, "d2h_bytes": int(d2h)})
    return out


def map_and_classify(ctx: capi.Context, index: capi.Index, *, reads=None, dev_ptr=None, host_ptr=None, offsets=None, read_len=None,
                     contig_len: np.ndarray, contig_taxon: np.ndarray, n_taxa: int, perc_identity: float = 80.0,
                     min_read_len: int = 1000, em_max_iter: int = 0, stats: dict | None = None, staged_slot: int | None = None):
    """One pass of the hot path over one batch of reads.  Returns per-mapping arrays + EM result.

    mm_map_batch (K1,K3-K5) -> mm_map_fetch_mappings (accepted mappings compacted on the device, identities on the host
    through glibc) -> mm_mapq_batch (K6) -> mm_nloc_batch (host) -> mm_em_run (K7/K8)."""
    import time
    t0 = time.perf_counter()
    res = capi.map_reads(ctx, index, reads, perc_identity, min_read_len, dev_ptr=dev_ptr, host_ptr=host_ptr, offsets=offsets, fetch=False,
                         staged_slot=staged_slot)
    t1 = time.perf_counter()
    if stats is not None:
        stats["map"] = res["stats"]
    n_reads = res["_n"]
    m = capi.fetch_mappings(ctx, res["summary"]["n_mappings"])
    t2 = time.perf_counter()
    if read_len is None:
        read_len = np.array([len(r) for r in reads], np.int32) if reads is not None else np.diff(offsets).astype(np.int32)
    wall = {"map_call": (t1 - t0) * 1e3, "fetch_mappings": (t2 - t1) * 1e3}
    out = classify_mappings(ctx, m, n_reads, read_len, k=index.k, contig_len=contig_len, contig_taxon=contig_taxon, n_taxa=n_taxa,
                            em_max_iter=em_max_iter, wall=wall)
    if stats is not None:       # wall-clock split of one step (ms): the C-ABI calls and the glue between them
        stats["wall_ms"] = wall
    out["summary"] = res["summary"]
    out["gpu_ms"] += res["gpu_ms"]; out["launches"] += res["launches"]; out["d2h_bytes"] += m["d2h_bytes"]
    return out


_MAPPING_KEYS = ("read", "seq", "pos", "shared", "sketch", "strand", "identity", "identity_parsed")


def merge_shard_mappings(parts: list) -> dict:
    """Mappings of the same reads against several contig-range shards (each in read order, global contig ids), given in
    shard order -> one list in the reference's order: per read, the shards' mappings concatenated in shard order, which
    is (seqId, wpos) order because shards are consecutive contig ranges (computeMap.hpp:352, mapWrap.h:128-132)."""
    cat = {k_: np.concatenate([p[k_] for p in parts]) if parts else np.zeros(0) for k_ in _MAPPING_KEYS}
    order = np.argsort(cat["read"], kind="stable")
    return {k_: np.ascontiguousarray(v[order]) for k_, v in cat.items()}


def map_and_classify_sharded(ctx: capi.Context, shards: list, *, reads=None, dev_ptr=None, host_ptr=None, offsets=None, read_len=None,
                             contig_len: np.ndarray, contig_taxon: np.ndarray, n_taxa: int, perc_identity: float = 80.0,
                             min_read_len: int = 1000, em_max_iter: int = 0, exchange=None, read_range=None, stats: dict | None = None):
    """The same pass against a reference split into contig-range shards (capi.Index objects with .first_contig set).

    Every shard maps ALL reads.  Single process: `shards` are walked one after the other (the --maxmemory analogue).
    Multi-GPU: each rank passes its own shard(s) and `exchange`, a callable that all-gathers a picklable object over the
    ranks and returns the list in rank order (e.g. torch.distributed.all_gather_object); `read_range = (lo, hi)` is the
    block of reads this rank finalises (mapping quality needs a read's mappings from all shards, mapWrap.h:226-278); the
    EM taxon sums are all-reduced inside mm_em_run."""
    import time
    gpu_ms = 0.0; launches = 0; d2h = 0; parts = []; summary = None
    t0 = time.perf_counter()
    for ix in shards:
        res = capi.map_reads(ctx, ix, reads, perc_identity, min_read_len, dev_ptr=dev_ptr, host_ptr=host_ptr, offsets=offsets, fetch=False)
        gpu_ms += res["gpu_ms"]; launches += res["launches"]
        if stats is not None:
            stats["map"] = res["stats"]
        m = capi.fetch_mappings(ctx, res["summary"]["n_mappings"])
        d2h += m.pop("d2h_bytes")
        m["seq"] = m["seq"] + np.int32(getattr(ix, "first_contig", 0))
        parts.append(m)
        if summary is None:
            summary = dict(res["summary"])
        else:
            for k_ in ("n_candidates", "n_mappings"):
                summary[k_] += res["summary"][k_]
        n_reads = res["_n"]
    t1 = time.perf_counter()
    if exchange is not None:
        parts = [p for rank_parts in exchange(parts) for p in rank_parts]
    m = merge_shard_mappings(parts)
    if read_range is not None:
        lo, hi = read_range
        keep = (m["read"] >= lo) & (m["read"] < hi)
        m = {k_: np.ascontiguousarray(v[keep]) for k_, v in m.items()}
    t2 = time.perf_counter()
    if read_len is None:
        read_len = np.array([len(r) for r in reads], np.int32) if reads is not None else np.diff(offsets).astype(np.int32)
    wall = {"map_calls": (t1 - t0) * 1e3, "exchange_merge": (t2 - t1) * 1e3}
    out = classify_mappings(ctx, m, n_reads, read_len, k=shards[0].k, contig_len=contig_len, contig_taxon=contig_taxon, n_taxa=n_taxa,
                            em_max_iter=em_max_iter, wall=wall)
    if stats is not None:
        stats["wall_ms"] = wall
    summary["n_mappings_all_shards"] = int(sum(len(p["read"]) for p in parts))
    out["summary"] = summary
    out["gpu_ms"] += gpu_ms; out["launches"] += launches; out["d2h_bytes"] += d2h
    return out


_TAXON_CACHE: dict = {}


def _nloc(tax, m_seq, m_read, L, contig_len, contig_taxon, n_taxa):
    """numpy restatement of mm_nloc_batch (kept as the cross-check in tests/test_host_stats.py):
    sum over the taxon's contigs of (len >= L ? len-L+1 : [contig seen among this read's mappings])."""
    key = (id(contig_len), id(contig_taxon))
    if key not in _TAXON_CACHE:
        order = np.lexsort((contig_len, contig_taxon))
        lens = contig_len[order].astype(np.int64); tx = contig_taxon[order].astype(np.int64)
        start = np.searchsorted(tx, np.arange(n_taxa + 1))
        csum = np.concatenate([[0], np.cumsum(lens)])
        big = int(lens.max()) + 2 if len(lens) else 2
        tsum = np.bincount(contig_taxon, weights=contig_len.astype(np.float64), minlength=n_taxa).astype(np.int64)
        tcnt = np.bincount(contig_taxon, minlength=n_taxa).astype(np.int64)
        _TAXON_CACHE.clear(); _TAXON_CACHE[key] = (lens, start, csum, tx * big + lens, big, tsum, tcnt, int(lens.min()) if len(lens) else 0)
    lens, start, csum, skey, big, tsum, tcnt, min_len = _TAXON_CACHE[key]
    if len(L) and int(L.max()) <= min_len:          # every contig is at least as long as every read: closed form
        return (tsum[tax] - tcnt[tax] * (L - 1)).astype(np.float64)
    # contigs of taxon t are lens[start[t]:start[t+1]] ascending; p = first of them with len >= L: one global
    # searchsorted over (taxon, len) keys
    tax = tax.astype(np.int64)
    p = np.searchsorted(skey, tax * big + np.minimum(L, big - 1), side="left")
    p = np.clip(p, start[tax], start[tax + 1])
    n_big = start[tax + 1] - p
    big = (csum[start[tax + 1]] - csum[p]) - n_big * (L - 1)
    # short contigs (len < L) count once if this read maps to them; distinct (read, contig) pairs of the taxon
    short = contig_len[m_seq] < L
    seen = np.zeros(len(tax), np.int64)
    if short.any():
        idx = np.nonzero(short)[0]
        keyrc = m_read[idx].astype(np.int64) * (int(contig_len.shape[0]) + 1) + m_seq[idx]
        _, first = np.unique(keyrc, return_index=True)
        uniq_idx = idx[first]
        keyrt = m_read[uniq_idx].astype(np.int64) * (n_taxa + 1) + tax[uniq_idx]
        u, cnt = np.unique(keyrt, return_counts=True)
        allkey = m_read.astype(np.int64) * (n_taxa + 1) + tax
        pos = np.searchsorted(u, allkey)
        pos = np.minimum(pos, len(u) - 1)
        seen = np.where(u[pos] == allkey, cnt[pos], 0)
    return (big + seen).astype(np.float64)
