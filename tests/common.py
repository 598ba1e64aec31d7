"""Parity assertions shared by the host-emulation tests (CPU) and the GPU tests: every function takes a
metamaps_b200.capi.Context and compares what comes out of the C ABI with the oracle / golden fixtures."""
import numpy as np

from metamaps_b200 import capi, synth

ODD_PARAMS = ((16, 13), (16, 1), (11, 5), (8, 40), (5, 200))


def check_sketch_vs_golden(ctx, golden):
    odd = [s.encode() for s in golden["odd_seqs"]]
    for (k, w) in ODD_PARAMS:
        counts, hs, wp, st = ctx.sketch(odd, k, w)
        for i in range(len(odd)):
            a, b = counts[i], counts[i + 1]
            assert np.array_equal(hs[a:b], golden[f"odd{i}_k{k}_w{w}_hash"]), (i, k, w)
            assert np.array_equal(wp[a:b], golden[f"odd{i}_k{k}_w{w}_wpos"]), (i, k, w)
            assert np.array_equal(st[a:b], golden[f"odd{i}_k{k}_w{w}_strand"]), (i, k, w)


def check_sketch_vs_oracle(ctx, oracle, seqs, k, w):
    counts, hs, wp, st = ctx.sketch(seqs, k, w)
    assert counts[0] == 0 and counts[-1] == len(hs)
    for i, s in enumerate(seqs):
        oh, ow, os_ = oracle.minimizers(s, k, w)
        a, b = counts[i], counts[i + 1]
        assert np.array_equal(hs[a:b], oh), f"hash stream differs for sequence {i} (k={k}, w={w})"
        assert np.array_equal(wp[a:b], ow), f"wpos stream differs for sequence {i} (k={k}, w={w})"
        assert np.array_equal(st[a:b], os_), f"strand stream differs for sequence {i} (k={k}, w={w})"


def blockmin_stress_seqs(seed=3):
    """Sequences for K1's block-minimum path: random bases with runs the path must hand back to the deque kernel
    ((AT)n = k-mers with hashFwd == hashBwd, commonFunc.hpp:130; N runs and lower case, :44-51,106) placed on and
    around chunk boundaries (multiples of 32 .. 256 k-mer positions), homopolymers (equal hashes: newest wins, :144),
    and lengths around w, k and the chunk size."""
    rng = np.random.default_rng(seed)
    def rnd(n): return bytearray(rng.choice(list(b"ACGT"), size=n).astype(np.uint8).tobytes())
    seqs = []
    s = rnd(3000)
    for at, ln in ((96, 40), (250, 17), (511, 2), (640, 64), (1020, 30), (1279, 18), (2040, 200)):
        s[at:at + ln] = (b"AT" * ln)[:ln]
    seqs.append(bytes(s))
    s = rnd(2500)
    for at, ln in ((30, 1), (127, 3), (256, 1), (300, 50), (1023, 2), (1500, 400)):
        s[at:at + ln] = b"N" * ln
    s[700:900] = bytes(s[700:900]).lower()
    seqs.append(bytes(s))
    s = rnd(1500); s[200:420] = b"A" * 220; s[600:700] = b"ACGT" * 25; s[900:1100] = b"GC" * 100
    seqs.append(bytes(s))
    for n in (15, 16, 17, 31, 32, 33, 47, 48, 127, 128, 129, 143, 144, 145, 159, 160, 271, 272, 273):
        seqs.append(bytes(rnd(n)))
    seqs.append(b"AT" * 400)                      # every k-mer skipped
    seqs.append(bytes(rnd(20)) + b"AT" * 300 + bytes(rnd(500)))
    seqs.append(bytes(rnd(9000)))
    return seqs


BLOCKMIN_PARAMS = ((16, 16), (16, 13), (16, 2), (16, 32), (16, 31), (12, 17), (7, 5), (16, 33))


def build_index(ctx, contigs, k, w, batches=2):
    ix = capi.Index(ctx, k, w)
    step = max(1, (len(contigs) + batches - 1) // batches)
    for i in range(0, len(contigs), step):
        ix.add(contigs[i:i + step])
    ix.finalize()
    return ix


def check_index_vs_golden(ctx, golden, small):
    contigs = [synth.codes_to_ascii(c) for c in small["db"].contig_codes]
    ix = build_index(ctx, contigs, int(golden["k"]), int(golden["w"]))
    hs, sq, wp, st = ix.fetch()
    assert np.array_equal(hs, golden["index_hash"])
    assert np.array_equal(sq, golden["index_seq"])
    assert np.array_equal(wp, golden["index_wpos"])
    assert np.array_equal(st, golden["index_strand"])
    s = ix.stats()
    assert s["n_unique"] == int(golden["index_unique"])
    assert s["freq_threshold"] == int(golden["freq_threshold"])
    # lookup table: every present hash reports its multiplicity, absent hashes 0
    uniq, cnt = np.unique(golden["index_hash"], return_counts=True)
    assert np.array_equal(ix.lookup(uniq), cnt.astype(np.int32))
    absent = np.setdiff1d(np.arange(1, 200000, 7, dtype=np.uint32), uniq)
    assert not ix.lookup(absent).any()
    return ix


def check_map_vs_golden(ctx, golden, small, ix=None):
    if ix is None:
        contigs = [synth.codes_to_ascii(c) for c in small["db"].contig_codes]
        ix = build_index(ctx, contigs, int(golden["k"]), int(golden["w"]))
    reads = [synth.codes_to_ascii(r) for r in small["reads"]]
    res = capi.map_reads(ctx, ix, reads, 80.0, 1000, fetch_sketch=True)
    assert np.array_equal(res["s"], golden["read_s"])
    assert np.array_equal(res["minimumHits"], golden["read_minhits"])
    assert np.array_equal(res["cand_off"], golden["cand_off"])
    assert np.array_equal(res["q_off"], golden["q_off"])
    assert np.array_equal(res["q_hash"], golden["q_hash"])
    assert np.array_equal(res["q_strand"], golden["q_strand"])
    for key in ("seq", "start", "end", "shared", "valid", "votes", "optStart", "optEnd"):
        assert np.array_equal(res[key], golden["cand_" + key]), key
    v = golden["cand_valid"].astype(bool)       # position is undefined in the reference when nothing was shared
    assert np.array_equal(res["pos"][v], golden["cand_pos"][v])
    return res


def check_map_vs_oracle(ctx, oracle, contigs, reads, k, w, pi=80.0, min_len=1000, batches=2):
    ho = oracle.index_build(contigs, k, w)
    ix = build_index(ctx, contigs, k, w, batches)
    a = oracle.index_get(ho); b = ix.fetch()
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    assert ix.stats()["freq_threshold"] == oracle.index_freq_threshold(ho)
    res = capi.map_reads(ctx, ix, reads, pi, min_len, fetch_sketch=True)
    n_acc = 0
    for i, s in enumerate(reads):
        c0, c1 = res["cand_off"][i], res["cand_off"][i + 1]
        if len(s) < max(w, k, min_len):
            assert res["s"][i] == 0 and c0 == c1
            continue
        mo = oracle.map_read(ho, s, pi)
        assert mo["s"] == res["s"][i], i
        if mo["s"] == 0:
            assert c0 == c1
            continue
        assert mo["minimumHits"] == res["minimumHits"][i], i
        qh, qw, qs = oracle.read_sketch(s, k, w)
        q0, q1 = res["q_off"][i], res["q_off"][i + 1]
        assert np.array_equal(qh, res["q_hash"][q0:q1]), i
        assert np.array_equal(qs, res["q_strand"][q0:q1]), f"read {i}: surviving duplicate differs from std::sort+unique"
        for key in ("seq", "start", "end", "shared", "valid", "votes", "optStart", "optEnd"):
            assert np.array_equal(mo[key], res[key][c0:c1]), (i, key, mo[key], res[key][c0:c1])
        v = mo["valid"].astype(bool)
        assert np.array_equal(mo["pos"][v], res["pos"][c0:c1][v]), i
        for j in range(c1 - c0):
            nuc, up = oracle.identity(int(mo["shared"][j]), mo["s"], k)
            acc = int(bool(mo["valid"][j]) and up >= pi)
            assert acc == res["accepted"][c0 + j], (i, j)
            n_acc += acc
    assert n_acc == res["summary"]["n_mappings"]
    oracle.index_free(ho)
    return res


def random_mapq_case(seed, nr=400):
    rng = np.random.default_rng(seed)
    cnt = rng.integers(1, 12, nr); off = np.zeros(nr + 1, np.int64); off[1:] = np.cumsum(cnt); M = int(off[-1])
    sk = np.repeat(rng.integers(200, 1500, nr), cnt).astype(np.int32)
    sh = (sk * rng.uniform(0.02, 0.3, M)).astype(np.int32)
    ident = rng.uniform(0.78, 0.95, M).round(6)
    rl = rng.integers(1000, 20000, nr).astype(np.int32)
    return ident, sh, sk, rl, off


def check_mapq_vs_oracle(ctx, oracle, seed=3):
    ident, sh, sk, rl, off = random_mapq_case(seed)
    q, st = ctx.mapq(ident, sh, sk, rl, off, 16)
    for r in range(len(rl)):
        a, b = off[r], off[r + 1]
        rc, qo = oracle.mapq(ident[a:b], sh[a:b], sk[a:b], int(rl[r]), 16)
        if rc == 0:
            assert st[r] == 0
            assert np.abs(qo - q[a:b]).max() <= 1e-6        # north_star tolerance for mapping qualities
            assert abs(q[a:b].sum() - 1) < 1e-9
        else:
            assert st[r] == 1


def random_em_case(seed, nr=3000, T=40, maxc=12):
    rng = np.random.default_rng(seed)
    cnt = rng.integers(1, maxc, nr); off = np.zeros(nr + 1, np.int64); off[1:] = np.cumsum(cnt); M = int(off[-1])
    ab = rng.dirichlet(np.ones(T) * 0.3)
    tax = rng.choice(T, size=M, p=ab).astype(np.int32)
    mq = np.concatenate([(lambda v: v / v.sum())(rng.random(c)) for c in cnt])
    nloc = rng.integers(1000, 5_000_000, M).astype(np.float64)
    return tax, mq, nloc, off, T


def check_em_vs_oracle(ctx, oracle, seed=5, max_iter=0, **kw):
    tax, mq, nloc, off, T = random_em_case(seed, **kw)
    a = oracle.em(tax, mq, nloc, off, T, max_iter)
    b = ctx.em(tax, mq, nloc, off, T, max_iter)
    assert a["iters"] == b["iters"]
    assert np.abs(a["f"] - b["f"]).max() <= 1e-6             # north_star tolerance for EM frequencies
    assert abs(b["f"].sum() - 1) < 1e-9
    assert np.abs(a["posterior"] - b["posterior"]).max() <= 1e-6
    assert np.abs(a["ll"] - b["ll"]).max() <= 1e-6 * np.abs(a["ll"]).max()
    # read -> taxon assignment: bit-exact wherever the oracle's best posterior is not a near-tie
    diff = a["best"] != b["best"]
    for r in np.nonzero(diff)[0]:
        pa = a["posterior"][a["best"][r]]; pb = a["posterior"][b["best"][r]]
        assert abs(pa - pb) < 1e-12, r
    return a, b


def check_pipeline_vs_reference_files(ctx, small, golden_dir):
    """mapDirectly + classify on arrays vs the text files the unmodified reference wrote for the same inputs."""
    import os
    import re
    from metamaps_b200 import pipeline
    db = small["db"]
    contigs = [synth.codes_to_ascii(c) for c in db.contig_codes]
    reads = [synth.codes_to_ascii(r) for r in small["reads"]]
    ix = build_index(ctx, contigs, 16, 13)
    taxa = sorted(set(db.contig_taxon))
    tidx = {t: i for i, t in enumerate(taxa)}
    contig_taxon = np.array([tidx[t] for t in db.contig_taxon], np.int32)
    contig_len = np.array([len(c) for c in db.contig_codes], np.int64)
    out = pipeline.map_and_classify(ctx, ix, reads=reads, contig_len=contig_len, contig_taxon=contig_taxon, n_taxa=len(taxa))
    ref_lines = open(os.path.join(golden_dir, "ref_small", "ref")).read().splitlines()
    assert len(ref_lines) == len(out["read"])
    for m, line in enumerate(ref_lines):
        f = line.split(" ")
        r = out["read"][m]
        assert f[0] == small["names"][r] and int(f[1]) == len(reads[r])
        assert f[4] == ("+" if out["strand"][m] > 0 else "-")
        assert f[5] == db.contig_names[out["seq"][m]] and int(f[6]) == contig_len[out["seq"][m]]
        assert int(f[7]) == out["pos"][m] and int(f[8]) == out["pos"][m] + len(reads[r]) - 1      # mapping coordinates: bit-exact
        assert f[9] == "%g" % out["identity"][m]
        assert int(f[10]) == out["shared"][m] and int(f[11]) == out["sketch"][m]                  # MinHash counts: bit-exact
        assert abs(float(f[13]) - out["mapq"][m]) <= 1e-6 + 2e-6 * float(f[13])                    # 6 significant digits in the file
    meta = dict(l.split() for l in open(os.path.join(golden_dir, "ref_small", "ref.meta")))
    assert int(meta["ReadsTooShort"]) == out["summary"]["n_too_short"]
    assert int(meta["ReadsMapped"]) == len(out["mapped_reads"])
    em = out["em"]
    log = open(os.path.join(golden_dir, "ref_small", "classify.log")).read()
    assert em["iters"] == len(re.findall(r"^EM round", log, re.M))
    for m, line in enumerate(open(os.path.join(golden_dir, "ref_small", "ref.EM")).read().splitlines()):
        assert abs(float(line.split(" ")[13]) - em["posterior"][m]) <= 1e-6
    r2t = dict(l.rstrip("\n").split("\t") for l in open(os.path.join(golden_dir, "ref_small", "ref.EM.reads2Taxon")))
    for i, r in enumerate(out["mapped_reads"]):
        assert r2t[small["names"][r]] == taxa[out["taxon"][em["best"][i]]]                          # read -> taxon: exact
    for line in open(os.path.join(golden_dir, "ref_small", "ref.EM.WIMP")):
        c = line.rstrip("\n").split("\t")
        if c[0] == "definedGenomes" and c[1] in tidx:
            assert abs(float(c[4]) - em["f"][tidx[c[1]]]) <= 1e-6
    return out


def fallback_workloads():
    """Inputs that push the device fast paths into their fallbacks (each returns contigs, reads, k, w, min_len)."""
    rng = np.random.default_rng(123)
    out = {}
    # (a) more contigs than the 32768 per-read contig bins of the filtered gather (bins are then hashed)
    big = [rng.integers(0, 4, 30000, dtype=np.uint8) for _ in range(3)]
    tiny = [rng.integers(0, 4, 40, dtype=np.uint8) for _ in range(33000)]
    contigs = tiny[:20000] + [big[0]] + tiny[20000:] + big[1:]
    db = synth.SynthDB([f"C{i}|kraken:taxid|{i}|x" for i in range(len(contigs))], [str(i) for i in range(len(contigs))], contigs)
    sub = synth.SynthDB(["a", "b", "c"], ["1", "2", "3"], big)
    _, reads, _ = synth.make_reads(sub, 7, 40, 2500, err=0.08)
    out["many_contigs"] = ([synth.codes_to_ascii(c) for c in db.contig_codes], [synth.codes_to_ascii(r) for r in reads], 16, 8, 1000)
    # (b) one very long read: sketch larger than a sweep slice / the classify staging buffer -> global-memory paths
    g = rng.integers(0, 4, 900_000, dtype=np.uint8)
    sub = synth.SynthDB(["g"], ["1"], [g])
    _, rl, _ = synth.make_reads(sub, 9, 3, 600_000, err=0.05)
    _, rs, _ = synth.make_reads(sub, 10, 20, 3000, err=0.08)
    out["long_read"] = ([synth.codes_to_ascii(g)], [synth.codes_to_ascii(r) for r in rl + rs], 16, 5, 1000)
    # (c) low-complexity read: a handful of distinct minimizers, so hundreds of reference-only hashes share one gap
    #     (8-bit gap counters overflow -> the candidate is replayed with 16-bit counters in global memory)
    unit = rng.integers(0, 4, 23, dtype=np.uint8)
    rep = np.tile(unit, 400)                                   # 9.2 kb tandem repeat
    flankL = rng.integers(0, 4, 20000, dtype=np.uint8); flankR = rng.integers(0, 4, 20000, dtype=np.uint8)
    c0 = np.concatenate([flankL, rep, flankR])
    reads = [np.concatenate([flankL[-1500:], rep[:3000]]), rep[100:5100].copy(), np.concatenate([rep[-2500:], flankR[:2500]])]
    out["low_complexity"] = ([synth.codes_to_ascii(c0)], [synth.codes_to_ascii(r) for r in reads], 16, 6, 1000)
    return out


def sharded_case(seed=5, n_contigs=6, L=200_000, w=10, n_reads=60):
    """A reference whose global occurrence threshold is finite (70) although no half of it reaches it: five 25-base units
    (one full window of w = 10 k-mers, so their smallest k-mer is a minimizer in every copy) planted 90/70/55/41/33 times
    over all contigs.  Returns (db, contig ASCII list, read ASCII list, w, contig_taxon, contig_len, n_taxa)."""
    rng = np.random.default_rng(seed)
    contigs = [rng.integers(0, 4, L, dtype=np.uint8) for _ in range(n_contigs)]
    for copies in (90, 70, 55, 41, 33):
        unit = rng.integers(0, 4, 25, dtype=np.uint8)
        for _ in range(copies):
            ci = int(rng.integers(0, n_contigs)); pos = int(rng.integers(0, L - 50))
            contigs[ci][pos:pos + 25] = unit
    names = [f"C{i}|kraken:taxid|{i + 1}|x" for i in range(n_contigs)]
    db = synth.SynthDB(names, [str(i + 1) for i in range(n_contigs)], contigs)
    _, reads, _ = synth.make_reads(db, seed + 1, n_reads, 3000, err=0.08)
    contig_taxon = np.arange(n_contigs, dtype=np.int32)
    contig_len = np.full(n_contigs, L, np.int64)
    return db, [synth.codes_to_ascii(c) for c in contigs], [synth.codes_to_ascii(r) for r in reads], w, contig_taxon, contig_len, n_contigs


MAPPING_KEYS = ("read", "seq", "pos", "shared", "sketch", "strand", "identity", "mapq")


def check_shard_walk_equals_full(ctx, contigs, reads, k, w, contig_taxon, contig_len, n_taxa, cuts, min_len=1000):
    """Single process: the reference cut into contig-range shards walked one after the other (per-shard thresholds, the
    reference's --maxmemory behaviour) gives the unsharded result whenever no hash is over-frequent."""
    from metamaps_b200 import capi, pipeline
    full = build_index(ctx, contigs, k, w)
    a = pipeline.map_and_classify(ctx, full, reads=reads, contig_len=contig_len, contig_taxon=contig_taxon, n_taxa=n_taxa, min_read_len=min_len)
    shards = []
    bounds = [0] + list(cuts) + [len(contigs)]
    for c0, c1 in zip(bounds[:-1], bounds[1:]):
        ix = capi.Index(ctx, k, w); ix.set_shard(c0, keep_counts=False); ix.add(contigs[c0:c1]); ix.finalize(); shards.append(ix)
    b = pipeline.map_and_classify_sharded(ctx, shards, reads=reads, contig_len=contig_len, contig_taxon=contig_taxon, n_taxa=n_taxa, min_read_len=min_len)
    assert len(a["read"]) > 0
    for key in MAPPING_KEYS:
        assert np.array_equal(a[key], b[key]), key
    assert a["em"]["iters"] == b["em"]["iters"]
    # taxon sums are accumulated with atomics on the device: the order of the additions, hence the last bits, may differ
    assert np.abs(a["em"]["f"] - b["em"]["f"]).max() <= 1e-12 and np.array_equal(a["em"]["best"], b["em"]["best"])
    return a


def check_staged_equals_direct(ctx, contigs, reads, k, w):
    """mm_stage_reads_async + mm_map_batch_staged (both slots, staged ahead) give what mm_map_batch gives."""
    from metamaps_b200 import capi
    ix = build_index(ctx, contigs, k, w)
    half = len(reads) // 2
    batches = [reads[:half], reads[half:]]
    direct = [capi.map_reads(ctx, ix, b, 80.0, 1000) for b in batches]
    bufs = []
    for b in batches:
        data = np.frombuffer(b"".join(b), np.uint8).copy()
        off = np.zeros(len(b) + 1, np.int64); off[1:] = np.cumsum([len(x) for x in b])
        bufs.append((data, off))
    ctx.stage_reads(0, bufs[0][0].ctypes.data, bufs[0][1])
    ctx.stage_reads(1, bufs[1][0].ctypes.data, bufs[1][1])          # staged before batch 0 is mapped
    for slot in (0, 1):
        got = capi.map_reads(ctx, ix, None, 80.0, 1000, offsets=bufs[slot][1], staged_slot=slot)
        for key in ("s", "cand_off", "seq", "pos", "shared", "votes", "accepted", "valid"):
            assert np.array_equal(got[key], direct[slot][key]), (slot, key)


def check_index_save_load(ctx, contigs, reads, k, w, path):
    """mm_index_save -> mm_index_load: identical arrays and identical mapping results."""
    from metamaps_b200 import capi
    ix = build_index(ctx, contigs, k, w)
    ix.save(path)
    ix2 = capi.Index.load(ctx, path)
    sa, sb = ix.stats(), ix2.stats()
    sa.pop("device_bytes"); sb.pop("device_bytes")          # allocation sizes, not content
    assert (ix2.k, ix2.w) == (k, w) and sa == sb, (sa, sb)
    for a, b in zip(ix.fetch(), ix2.fetch()):
        assert np.array_equal(a, b)
    probe = np.concatenate([ix.fetch()[0][:2000], np.arange(1000, dtype=np.uint32) * 7919])
    assert np.array_equal(ix.lookup(probe), ix2.lookup(probe))
    r1 = capi.map_reads(ctx, ix, reads, 80.0, 1000); r2 = capi.map_reads(ctx, ix2, reads, 80.0, 1000)
    for key in ("s", "cand_off", "seq", "start", "end", "pos", "shared", "votes", "accepted", "valid", "optStart", "optEnd"):
        assert np.array_equal(r1[key], r2[key]), key
    assert r1["summary"]["n_mappings"] > 0
    import os
    import pytest
    size = os.path.getsize(path)
    with open(path, "r+b") as f:          # a truncated file is refused before anything is allocated from its header
        f.truncate(size - 64)
    with pytest.raises(capi.MMError, match="file size"):
        capi.Index.load(ctx, path)
    with open(path, "r+b") as f:          # a damaged file is refused, not loaded
        f.write(b"XXXX")
    with pytest.raises(capi.MMError):
        capi.Index.load(ctx, path)
    # an index without minimizers (every contig shorter than w / k) saves, loads and maps: nothing is found
    empty = capi.Index(ctx, k, w); empty.add([b"ACGTACG", b"TTGA"]); empty.finalize()
    assert empty.stats()["n_minimizers"] == 0 and empty.stats()["n_contigs"] == 2
    empty.save(path + ".empty")
    e2 = capi.Index.load(ctx, path + ".empty")
    assert e2.stats()["n_minimizers"] == 0 and e2.stats()["n_contigs"] == 2
    for ie in (empty, e2):
        r0 = capi.map_reads(ctx, ie, reads[:5], 80.0, 1000)
        assert r0["summary"]["n_candidates"] == 0 and r0["summary"]["n_mappings"] == 0


def check_api_errors(ctx, tmp_path):
    """Misuse of the session-2 entry points fails loudly with an MM_E* code and a message, never silently."""
    import ctypes as C
    import pytest
    from metamaps_b200 import capi
    lib = ctx.lib
    contigs = [synth.codes_to_ascii(np.random.default_rng(1).integers(0, 4, 30_000, dtype=np.uint8))]
    ix = build_index(ctx, contigs, 16, 13)
    with pytest.raises(capi.MMError) as e:
        ix.set_shard(3)                                   # after finalize
    assert e.value.code == -22
    with pytest.raises(capi.MMError):
        ix.sync_threshold()                               # shard counts were not kept
    with pytest.raises(capi.MMError):
        capi.Index.load(ctx, str(tmp_path / "does_not_exist.0"))
    with pytest.raises(capi.MMError):
        ix.save(str(tmp_path / "no_such_dir" / "ix.0"))
    fresh = capi.Context(0, lib)
    p = capi.MapParams(80.0, 1000, 1, 0); s = capi.MapSummary()
    assert lib.mm_map_batch_staged(fresh.h, ix.h, 0, C.byref(p), C.byref(s)) == -22      # nothing staged in that slot
    assert b"staged" in lib.mm_last_error()
    fresh.close()
    reads = [contigs[0][1000:4000], contigs[0][8000:12000]]
    res = capi.map_reads(ctx, ix, reads, 80.0, 1000, fetch=False)
    assert res["summary"]["n_mappings"] >= 2
    got = C.c_int64(); buf = np.zeros(1, np.int32)
    rc = lib.mm_map_fetch_mappings(ctx.h, buf.ctypes.data, None, None, None, None, None, None, None, 1, C.byref(got))
    assert rc == -34 and got.value == res["summary"]["n_mappings"]                        # capacity too small, count still reported
    with pytest.raises(capi.MMError):
        capi.nloc_batch(lib, np.array([5], np.int32), np.array([0, 1], np.int64), np.array([100], np.int32), np.array([1000], np.int64),
                        np.array([0], np.int32), 1)                                       # contig id out of range
    with pytest.raises(capi.MMError):
        capi.nloc_batch(lib, np.array([0], np.int32), np.array([0, 1], np.int64), np.array([100], np.int32), np.array([1000], np.int64),
                        np.array([7], np.int32), 1)                                       # taxon out of range


def check_multi_batch_classify(ctx, small):
    """Two read batches appended to ONE classify table (mm_classify_next_batch) and one EM over both = the one-batch run;
    and the non---all filter inside the library (mm_map_params.report_all = 0) = reportReadMappings' identity >= best - 1.0."""
    from metamaps_b200 import capi, pipeline
    db = small["db"]
    contigs = [synth.codes_to_ascii(c) for c in db.contig_codes]
    reads = [synth.codes_to_ascii(r) for r in small["reads"]]
    taxa = sorted(set(db.contig_taxon)); tidx = {t: i for i, t in enumerate(taxa)}
    contig_taxon = np.array([tidx[t] for t in db.contig_taxon], np.int32)
    contig_len = np.array([len(c) for c in db.contig_codes], np.int64)
    ix = build_index(ctx, contigs, 16, 13)
    one = pipeline.map_and_classify(ctx, ix, reads=reads, contig_len=contig_len, contig_taxon=contig_taxon, n_taxa=len(taxa))
    one = {k_: (np.array(v) if isinstance(v, np.ndarray) else v) for k_, v in one.items()}
    one["em"] = {k_: (np.array(v) if isinstance(v, np.ndarray) else v) for k_, v in one["em"].items()}
    ctx.classify_setup(contig_len, contig_taxon, len(taxa))
    ctx.classify_begin()
    cut = len(reads) // 3
    for part in (reads[:cut], reads[cut:]):
        capi.map_reads(ctx, ix, part, 80.0, 1000, fetch=False)
        ctx.classify_add(0); ctx.classify_next_batch()
    cs = ctx.classify_run(0)
    two = ctx.classify_fetch(cs)
    for key in ("read", "seq", "pos", "shared", "sketch", "strand", "identity", "mapq", "taxon", "nloc", "mapped_reads", "read_off"):
        assert np.array_equal(one[key], two[key]), key
    assert cs["em_iters"] == one["em"]["iters"]
    assert np.abs(two["f"] - one["em"]["f"]).max() <= 1e-12 and np.array_equal(two["best"], one["em"]["best"])
    # em_max_iter < 0: everything but the EM
    ctx.classify_begin(); capi.map_reads(ctx, ix, reads, 80.0, 1000, fetch=False); ctx.classify_add(0)
    cs0 = ctx.classify_run(-1)
    assert cs0["em_iters"] == 0 and cs0["n_mappings"] == len(one["read"])
    # report_all = 0
    import ctypes as C
    p = capi.MapParams(80.0, 1000, 0, 0); s = capi.MapSummary()
    data, offs = capi._ascii_batch(reads)
    ctx._check(ctx.lib.mm_map_batch(ctx.h, ix.h, data, offs, len(reads), C.byref(p), C.byref(s)))
    ctx.classify_begin(); ctx.classify_add(0)
    top = ctx.classify_fetch(ctx.classify_run(-1))
    keep = np.zeros(len(one["read"]), bool)
    for g in range(len(one["mapped_reads"])):
        a, b = one["read_off"][g], one["read_off"][g + 1]
        best = np.float32(0)
        for m in range(a, b):
            if one["identity"][m] > best:
                best = one["identity"][m]
        keep[a:b] = one["identity"][a:b].astype(np.float64) >= float(best) - 1.0
    assert keep.sum() < len(keep) and s.n_mappings == keep.sum()
    for key in ("read", "seq", "pos", "shared", "identity"):
        assert np.array_equal(top[key], one[key][keep]), key


def check_prune_fuzz(make_ctx, monkeypatch, cases, before_pruned=None):
    """MM_SWEEP_PRUNE=1 against MM_SWEEP_PRUNE=0 on random repeat-rich references (a fresh context per run: the switch is read at creation).
    before_pruned(case): called before every pruned run (the emulation tier moves the pass's ranks with it)."""
    from metamaps_b200 import capi, synth
    rng = np.random.default_rng(20261017)
    swept = total = 0
    for case in range(cases):
        w = int(rng.choice([3, 5, 8, 13, 16, 24])); k = int(rng.choice([12, 14, 16]))
        L = int(rng.integers(20000, 90000))
        base = rng.integers(0, 4, L, dtype=np.uint8)
        for _ in range(int(rng.integers(0, 6))):
            u = int(rng.integers(50, 2500)); a = int(rng.integers(0, L - u)); b = int(rng.integers(0, L - u))
            seg = base[a:a + u].copy(); m = rng.random(u) < rng.choice([0.0, 0.01, 0.05]); seg[m] = (seg[m] + 1) % 4
            base[b:b + u] = seg
        contigs = [base]
        for _ in range(int(rng.integers(0, 3))):
            c = base.copy(); m = rng.random(L) < rng.choice([0.002, 0.01, 0.03]); c[m] = (c[m] + rng.integers(1, 4, int(m.sum()))) % 4
            a = int(rng.integers(0, L // 2)); contigs.append(c[a:a + int(rng.integers(L // 4, L // 2))])
        db = synth.SynthDB([f"C{i}|kraken:taxid|{i + 1}|x" for i in range(len(contigs))], [str(i + 1) for i in range(len(contigs))], contigs)
        _, reads, _ = synth.make_reads(db, int(rng.integers(1, 1 << 30)), int(rng.integers(10, 30)), int(rng.choice([1500, 3000, 6000])),
                                       err=float(rng.choice([0.0, 0.02, 0.08, 0.15])))
        asc = [synth.codes_to_ascii(c) for c in contigs]; rasc = [synth.codes_to_ascii(r) for r in reads]
        out = {}
        for flag in ("1", "0"):
            monkeypatch.setenv("MM_SWEEP_PRUNE", flag)
            if flag == "1" and before_pruned is not None:
                before_pruned(case)
            ctx = make_ctx()
            ix = build_index(ctx, asc, k, w)
            out[flag] = capi.map_reads(ctx, ix, rasc, 80.0, 1000)
            if flag == "1":
                st = ctx.last_map_stats(); swept += st["window_starts_swept"]; total += st["window_starts"]
            ix.close(); ctx.close()
        for key in ("seq", "start", "end", "shared", "valid", "votes", "optStart", "optEnd", "pos", "accepted"):
            assert np.array_equal(out["1"][key], out["0"][key]), (case, key, k, w)
    assert 0 < swept <= total and (before_pruned is not None or swept < 0.8 * total), (swept, total)
