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


def _classify_on_device(ctx: capi.Context, *, contig_len, contig_taxon, n_taxa: int, em_max_iter: int, wall: dict | None) -> dict:
    """mm_classify_run + mm_classify_fetch over the context's mapping table -> the result layout of this module."""
    import time
    t0 = time.perf_counter()
    cs = ctx.classify_run(em_max_iter)
    t1 = time.perf_counter()
    gpu_ms, launches = ctx.last_timing()
    r = ctx.classify_fetch(cs)
    t2 = time.perf_counter()
    if wall is not None:
        wall.update({"classify_call": (t1 - t0) * 1e3, "classify_fetch": (t2 - t1) * 1e3})
    em = None
    if cs["em_iters"] > 0:
        em = {"f": r["f"], "posterior": r["posterior"], "best": r["best"], "ll": r["ll"], "iters": cs["em_iters"]}
    out = {k_: r[k_] for k_ in ("read", "seq", "pos", "shared", "sketch", "strand", "identity", "identity_parsed", "mapq", "taxon", "nloc",
                                "mapped_reads", "read_off", "mapq_status")}
    out.update({"em": em, "gpu_ms": gpu_ms, "launches": launches, "d2h_bytes": r["d2h_bytes"], "classify": cs})
    return out


def map_and_classify(ctx: capi.Context, index: capi.Index, *, reads=None, dev_ptr=None, host_ptr=None, offsets=None, read_len=None,
                     contig_len: np.ndarray, contig_taxon: np.ndarray, n_taxa: int, perc_identity: float = 80.0,
                     min_read_len: int = 1000, em_max_iter: int = 0, stats: dict | None = None, staged_slot: int | None = None):
    """One pass of the hot path over one batch of reads.  Returns per-mapping arrays + EM result.

    mm_map_batch (K1,K3-K5) -> mm_classify_add_mappings -> mm_classify_run (identity, K6, nLoc, K7/K8: everything stays in
    HBM) -> mm_classify_fetch (one D2H of the finished arrays)."""
    import time
    t0 = time.perf_counter()
    ctx.classify_setup(contig_len, contig_taxon, n_taxa)
    res = capi.map_reads(ctx, index, reads, perc_identity, min_read_len, dev_ptr=dev_ptr, host_ptr=host_ptr, offsets=offsets, fetch=False,
                         staged_slot=staged_slot)
    t1 = time.perf_counter()
    if stats is not None:
        stats["map"] = res["stats"]
    ctx.classify_begin()
    ctx.classify_add(getattr(index, "first_contig", 0))
    wall = {"map_call": (t1 - t0) * 1e3}
    out = _classify_on_device(ctx, contig_len=contig_len, contig_taxon=contig_taxon, n_taxa=n_taxa, em_max_iter=em_max_iter, wall=wall)
    if stats is not None:       # wall-clock split of one step (ms): the C-ABI calls
        stats["wall_ms"] = wall
    out["summary"] = res["summary"]
    out["gpu_ms"] += res["gpu_ms"]; out["launches"] += res["launches"]
    return out


def map_and_classify_sharded(ctx: capi.Context, shards: list, *, reads=None, dev_ptr=None, host_ptr=None, offsets=None, read_len=None,
                             contig_len: np.ndarray, contig_taxon: np.ndarray, n_taxa: int, perc_identity: float = 80.0,
                             min_read_len: int = 1000, em_max_iter: int = 0, read_range=None, stats: dict | None = None, my_block=None):
    """The same pass against a reference split into contig-range shards (capi.Index objects with .first_contig set).

    Every shard maps ALL reads; each map call's accepted mappings are appended to the device-resident table with global contig
    ids (shard order = contig order, the reference's chunk order: mapWrap.h:128-132).  Single process: `shards` are walked one
    after the other (the --maxmemory analogue).  Multi-GPU (the context has a communicator, `read_range = (lo, hi)`): each rank
    passes its own shard(s); mm_classify_exchange all-gathers and merges the tables on the device and this rank finalises the
    reads [lo, hi) (mapping quality needs a read's mappings from all shards, mapWrap.h:226-278); the EM taxon sums are
    all-reduced inside mm_classify_run.
    `my_block = (dev_ptr, offsets)` (multi-GPU, one shard per rank): the rank holds only ITS block of the batch, sketches it once
    and the sketches are all-gathered (mm_map_batch_sharded_dev) instead of every rank re-sketching every read; read_range is
    then this rank's block."""
    import time
    gpu_ms = 0.0; launches = 0; summary = None
    t0 = time.perf_counter()
    ctx.classify_setup(contig_len, contig_taxon, n_taxa)
    ctx.classify_begin()
    n_all = 0
    for ix in shards:
        if my_block is not None:
            res = capi.map_reads_sharded(ctx, ix, my_block[0], my_block[1], perc_identity, min_read_len)
            read_range = (res["first_read"], res["first_read"] + len(my_block[1]) - 1)
        else:
            res = capi.map_reads(ctx, ix, reads, perc_identity, min_read_len, dev_ptr=dev_ptr, host_ptr=host_ptr, offsets=offsets, fetch=False)
        gpu_ms += res["gpu_ms"]; launches += res["launches"]
        if stats is not None:
            stats["map"] = res["stats"]
        n_all = ctx.classify_add(getattr(ix, "first_contig", 0))
        gpu_ms += ctx.last_timing()[0]; launches += ctx.last_timing()[1]
        if summary is None:
            summary = dict(res["summary"])
        else:
            for k_ in ("n_candidates", "n_mappings"):
                summary[k_] += res["summary"][k_]
    t1 = time.perf_counter()
    if read_range is not None and ctx.n_ranks_hint > 1:
        ctx.classify_exchange(read_range[0], read_range[1])
        gpu_ms += ctx.last_timing()[0]; launches += ctx.last_timing()[1]
    t2 = time.perf_counter()
    wall = {"map_calls": (t1 - t0) * 1e3, "exchange_merge": (t2 - t1) * 1e3}
    out = _classify_on_device(ctx, contig_len=contig_len, contig_taxon=contig_taxon, n_taxa=n_taxa, em_max_iter=em_max_iter, wall=wall)
    if stats is not None:
        stats["wall_ms"] = wall
    out["wall_ms"] = wall
    summary["n_mappings_this_rank_all_shards"] = int(n_all)
    out["summary"] = summary; out["read_range"] = read_range
    out["gpu_ms"] += gpu_ms; out["launches"] += launches
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


# ---- the reference's --maxmemory chunk chain (winSketch.hpp:284-329,452-495) for chunks built anywhere ------------------------
def walk_freq_threshold(cum_hist: dict, n_unique: int, prev_threshold: int) -> int:
    """computeFreqHist's walk (winSketch.hpp:462-488): cumulative histogram {occurrences: hashes}, THIS chunk's unique count,
    the threshold left by the previous chunk."""
    to_ignore = int(np.float32(np.float32(n_unique) * np.float32(0.001)) / np.float32(100))      # int64 * float -> float, / int -> float
    thr = prev_threshold; s_ = 0
    for k_ in sorted(cum_hist, reverse=True):
        s_ += cum_hist[k_]
        if s_ < to_ignore:
            thr = k_
        elif s_ == to_ignore:
            thr = k_
            break
        else:
            break
    return thr


def chunk_chain(own_hists: list, n_uniques: list):
    """own_hists[c] = [(occurrences, hashes)] of chunk c alone, in chunk order.  Returns per chunk (carry histogram before it,
    threshold before it, its own threshold) -- what mm_index_set_freq_carry needs for chunk c and what it must then report."""
    INT_MAX = 0x7fffffff
    cum: dict = {}; thr = INT_MAX; out = []
    for own, nu in zip(own_hists, n_uniques):
        carry = sorted(cum.items()); prev = thr
        if nu > 0:
            for k_, v in own:
                cum[k_] = cum.get(k_, 0) + v
            thr = walk_freq_threshold(cum, nu, prev)
        out.append((carry, prev, thr))
    return out


def map_and_classify_streamed(ctx: capi.Context, build_chunk, chunk_ids: list, n_chunks: int, all_gather, *, dev_ptr=None, reads=None, offsets=None,
                              contig_len, contig_taxon, n_taxa: int, perc_identity: float = 80.0, min_read_len: int = 1000, em_max_iter: int = 0,
                              read_range=None, stats: dict | None = None):
    """Config 5's shape: a reference too large to be resident is walked as `n_chunks` contig-range chunks, one resident at a time
    per GPU (`build_chunk(c)` -> capi.Index with .first_contig set, NOT carrying a threshold yet), this rank owning `chunk_ids`
    (ascending; ranks interleave: rank g of N owns g, g + N, ...).  Thresholds follow the reference's chunk chain (non-reset
    histogram): after each round of builds the ranks all-gather (`all_gather(obj) -> list`, e.g. torch.distributed
    all_gather_object; identity for one rank) the own histograms of that round's chunks and settle their carries.  Every chunk
    maps ALL reads; the tables are exchanged and merged by (read, contig) on the device at the end."""
    ctx.classify_setup(contig_len, contig_taxon, n_taxa)
    ctx.classify_begin()
    rounds = max(len(ids) for ids in all_gather(list(chunk_ids)))
    known: dict = {}                       # chunk id -> (own histogram, n_unique), learnt round by round
    gpu_ms = 0.0; launches = 0; summary = None; thresholds = {}
    for t in range(rounds):
        c = chunk_ids[t] if t < len(chunk_ids) else None
        ix = None; mine = None
        if c is not None:
            ix = build_chunk(c)
            mine = (c, ix.freq_hist()[0], ix.stats()["n_unique"])
        for item in all_gather(mine):
            if item is not None:
                known[item[0]] = (item[1], item[2])
        if c is not None:
            upto = sorted(k_ for k_ in known if k_ <= c)
            assert upto == list(range(c + 1)), "chunks must become known in order (interleaved ownership, one round at a time)"
            chain = chunk_chain([known[k_][0] for k_ in upto], [known[k_][1] for k_ in upto])
            carry, prev, thr = chain[c]
            ix.set_freq_carry(carry, prev)
            assert ix.stats()["freq_threshold"] == thr, (ix.stats()["freq_threshold"], thr)
            thresholds[c] = thr
            res = capi.map_reads(ctx, ix, reads, perc_identity, min_read_len, dev_ptr=dev_ptr, offsets=offsets, fetch=False)
            gpu_ms += res["gpu_ms"]; launches += res["launches"]
            if stats is not None:
                stats["map"] = res["stats"]
                acc = stats.setdefault("map_sum_ms", {})
                for k_, v in res["stats"].items():
                    if k_.endswith("_ms"):
                        acc[k_] = acc.get(k_, 0.0) + float(v)
            ctx.classify_add(ix.first_contig)
            if summary is None:
                summary = dict(res["summary"])
            else:
                for k_ in ("n_candidates", "n_mappings"):
                    summary[k_] += res["summary"][k_]
            ix.close()
    if read_range is not None and ctx.n_ranks_hint > 1:
        ctx.classify_exchange(read_range[0], read_range[1])
    out = _classify_on_device(ctx, contig_len=contig_len, contig_taxon=contig_taxon, n_taxa=n_taxa, em_max_iter=em_max_iter, wall=None)
    out["summary"] = summary; out["thresholds"] = thresholds
    out["gpu_ms"] += gpu_ms; out["launches"] += launches
    return out
