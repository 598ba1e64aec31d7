"""Parity tests proper: the CUDA path, called through the C ABI, against the oracle and the golden fixtures."""
import numpy as np
import pytest

from metamaps_b200 import capi, synth
from tests import common

pytestmark = pytest.mark.gpu


def test_sketch_golden(gpu_ctx, golden):
    common.check_sketch_vs_golden(gpu_ctx, golden)


def test_sketch_edge_cases(gpu_ctx, oracle):
    rng = np.random.default_rng(11)
    seqs = [b"", b"A", b"ACGTACGTACGTACG", b"ACGTACGTACGTACGT", b"N" * 300, b"ACGT" * 100 + b"N" * 64 + b"TTGACC" * 50,
            bytes(rng.choice(list(b"ACGT"), size=20001).astype(np.uint8)),
            bytes(rng.choice(list(b"acgtRYKM"), size=777).astype(np.uint8)), b"GATTACA" * 300]
    for (k, w) in ((16, 16), (16, 13), (16, 1), (15, 8), (9, 33), (4, 100), (16, 500), (16, 2000)):
        common.check_sketch_vs_oracle(gpu_ctx, oracle, seqs, k, w)


@pytest.mark.parametrize("env", [{}, {"MM_SKETCH_CH": "32"}, {"MM_SKETCH_CH": "48"}, {"MM_SKETCH_BLOCKMIN": "0"}])
def test_sketch_blockmin(gpu_ctx, oracle, monkeypatch, env):
    """K1 fast path (block prefix / suffix minima in shared memory) + bail list vs the oracle; see the emulation test."""
    for k_, v in env.items():
        monkeypatch.setenv(k_, v)
    seqs = common.blockmin_stress_seqs()
    for (k, w) in common.BLOCKMIN_PARAMS:
        common.check_sketch_vs_oracle(gpu_ctx, oracle, seqs, k, w)


def test_sketch_large_batch_properties(gpu_ctx, oracle):
    """2000 reads x ~8 kb: spot-check 40 against the oracle; window property on all (a minimizer stays the
    window minimum for at most w steps, so consecutive wpos differ by <= w, plus one step for every
    reverse-palindromic k-mer the reference skips, commonFunc.hpp:130)."""
    rng = np.random.default_rng(3)
    seqs = [bytes(rng.choice(list(b"ACGT"), size=int(rng.integers(1200, 16000))).astype(np.uint8)) for _ in range(2000)]
    counts, hs, wp, st = gpu_ctx.sketch(seqs, 16, 16)
    for i in rng.choice(len(seqs), 40, replace=False):
        oh, ow, os_ = oracle.minimizers(seqs[i], 16, 16)
        a, b = counts[i], counts[i + 1]
        assert np.array_equal(hs[a:b], oh) and np.array_equal(wp[a:b], ow) and np.array_equal(st[a:b], os_)
    for i in range(len(seqs)):
        w_ = wp[counts[i]:counts[i + 1]]
        assert w_[0] == 0 and (np.diff(w_) > 0).all() and (np.diff(w_) <= 16 + 2).all()
        assert w_[-1] <= len(seqs[i]) - 16 - 16 + 1


def test_index_golden(gpu_ctx, golden, small_workload):
    common.check_index_vs_golden(gpu_ctx, golden, small_workload)


def test_map_golden(gpu_ctx, golden, small_workload):
    res = common.check_map_vs_golden(gpu_ctx, golden, small_workload)
    assert res["summary"]["n_too_short"] == 4 and res["summary"]["n_reads_mapped"] == 194


def test_map_config1(gpu_ctx, oracle, tmp_path):
    """BASELINE config 1 (the reference's own CPU-runnable case): 10-genome mini DB, 1000 reads of 5 kb."""
    db, fa, fq, (names, reads, truth) = synth.config1(str(tmp_path), n_reads=1000)
    contigs = [synth.codes_to_ascii(c) for c in db.contig_codes]
    res = common.check_map_vs_oracle(gpu_ctx, oracle, contigs, [synth.codes_to_ascii(r) for r in reads], 16, 13)
    assert res["summary"]["n_reads_mapped"] == 967 and res["summary"]["n_too_short"] == 21   # reference .meta


def test_map_repetitive_reference(gpu_ctx, oracle):
    rng = np.random.default_rng(77)
    unit = rng.integers(0, 4, 700, dtype=np.uint8)
    parts = [rng.integers(0, 4, 5000, dtype=np.uint8), unit, unit, unit, rng.integers(0, 4, 3000, dtype=np.uint8), unit,
             rng.integers(0, 4, 4000, dtype=np.uint8)]
    c0 = np.concatenate(parts)
    c1 = np.concatenate([c0[2000:9000], rng.integers(0, 4, 2000, dtype=np.uint8), c0[2000:6000]])
    db = synth.SynthDB(["C0|kraken:taxid|1|x", "C1|kraken:taxid|2|x"], ["1", "2"], [c0, c1])
    _, reads, _ = synth.make_reads(db, 5, 60, 2500, err=0.06)
    contigs = [synth.codes_to_ascii(c) for c in db.contig_codes]
    common.check_map_vs_oracle(gpu_ctx, oracle, contigs, [synth.codes_to_ascii(r) for r in reads], 16, 5, 80.0, 1000)


def test_map_low_complexity_reads(gpu_ctx, oracle):
    rng = np.random.default_rng(9)
    genome = rng.integers(0, 4, 40000, dtype=np.uint8)
    pal = rng.integers(0, 4, 300, dtype=np.uint8)
    rc = (3 - pal[::-1]).astype(np.uint8)
    genome[10000:10300] = pal; genome[10300:10600] = rc; genome[12000:12300] = pal
    reads = [synth.codes_to_ascii(genome[9000:13500]), synth.codes_to_ascii(genome[9500:12800]),
             synth.codes_to_ascii((3 - genome[9000:13500][::-1]).astype(np.uint8))]
    common.check_map_vs_oracle(gpu_ctx, oracle, [synth.codes_to_ascii(genome)], reads, 16, 4, 80.0, 1000, batches=1)


def test_map_ragged_and_empty(gpu_ctx, oracle, small_workload):
    contigs = [synth.codes_to_ascii(c) for c in small_workload["db"].contig_codes]
    reads = [synth.codes_to_ascii(r) for r in small_workload["reads"][:12]]
    ragged = [b"", reads[0], b"ACGT", reads[1][:999], reads[2], b"N" * 1500, reads[3] + reads[4], reads[5][:1000]]
    common.check_map_vs_oracle(gpu_ctx, oracle, contigs + [b"ACGT", b""], ragged, 16, 13)
    ix = common.build_index(gpu_ctx, contigs, 16, 13)
    res = capi.map_reads(gpu_ctx, ix, [], 80.0, 1000)
    assert res["summary"]["n_reads"] == 0 and res["summary"]["n_candidates"] == 0


def test_map_is_deterministic_and_batch_invariant(gpu_ctx, small_workload):
    """Idempotence: mapping the same reads twice, or split into two batches, gives identical candidates."""
    contigs = [synth.codes_to_ascii(c) for c in small_workload["db"].contig_codes]
    reads = [synth.codes_to_ascii(r) for r in small_workload["reads"]]
    ix = common.build_index(gpu_ctx, contigs, 16, 13, batches=3)
    a = capi.map_reads(gpu_ctx, ix, reads)
    b = capi.map_reads(gpu_ctx, ix, reads)
    h1 = capi.map_reads(gpu_ctx, ix, reads[:77]); h2 = capi.map_reads(gpu_ctx, ix, reads[77:])
    for key in ("seq", "start", "end", "pos", "shared", "votes", "accepted"):
        assert np.array_equal(a[key], b[key])
        assert np.array_equal(a[key], np.concatenate([h1[key], h2[key]]))


def test_mapq(gpu_ctx, oracle):
    common.check_mapq_vs_oracle(gpu_ctx, oracle)


def test_em(gpu_ctx, oracle):
    common.check_em_vs_oracle(gpu_ctx, oracle, seed=5)                                            # 8 lanes per read
    common.check_em_vs_oracle(gpu_ctx, oracle, seed=6, max_iter=50, nr=500, T=7, maxc=30)         # 32 lanes per read, fixed round count
    common.check_em_vs_oracle(gpu_ctx, oracle, seed=8, nr=20000, T=300, maxc=40)
    common.check_em_vs_oracle(gpu_ctx, oracle, seed=9, nr=4000, T=50, maxc=4)                     # 4 lanes per read
    common.check_em_vs_oracle(gpu_ctx, oracle, seed=10, nr=3000, T=9000, maxc=12)                 # one 72 KB copy of the taxon sums per CTA
    common.check_em_vs_oracle(gpu_ctx, oracle, seed=11, nr=3000, T=40000, maxc=12)                # taxon sums too large for shared memory
    common.check_em_vs_oracle(gpu_ctx, oracle, seed=12, nr=300, T=200, maxc=400)                  # config-4-like: hundreds of mappings per read


def test_em_stopping_round_is_independent_of_the_host_check_interval(gpu_ctx, oracle, monkeypatch):
    """The stopping rule (fEM.h:636) runs on the device; rounds launched past it must not change anything."""
    tax, mq, nloc, off, T = common.random_em_case(5)
    ref = gpu_ctx.em(tax, mq, nloc, off, T, 0)
    for chk in ("1", "3", "16"):
        monkeypatch.setenv("MM_EM_CHECK", chk)
        r = gpu_ctx.em(tax, mq, nloc, off, T, 0)
        assert r["iters"] == ref["iters"] and np.abs(r["f"] - ref["f"]).max() <= 1e-12 and len(r["ll"]) == r["iters"]


def test_em_rejects_a_read_without_likelihood(gpu_ctx):
    """All mapping qualities of one read are 0: the reference asserts (fEM.h:357); the library must return an error, not loop."""
    tax, mq, nloc, off, T = common.random_em_case(5, nr=50)
    mq = mq.copy(); mq[off[7]:off[8]] = 0.0
    with pytest.raises(capi.MMError):
        gpu_ctx.em(tax, mq, nloc, off, T, 0)


def test_em_properties_large(gpu_ctx):
    """Size-independent properties at scale: f sums to 1, posteriors sum to 1 per read, log-likelihood is monotone."""
    tax, mq, nloc, off, T = common.random_em_case(21, nr=200_000, T=2000, maxc=20)
    r = gpu_ctx.em(tax, mq, nloc, off, T, 12)
    assert abs(r["f"].sum() - 1) < 1e-9
    sums = np.add.reduceat(r["posterior"], off[:-1])
    assert np.abs(sums - 1).max() < 1e-9
    assert (np.diff(r["ll"]) >= -1e-6).all()
    assert ((r["best"] >= off[:-1]) & (r["best"] < off[1:])).all()


def test_pipeline_matches_reference_files(gpu_ctx, small_workload):
    """End to end on arrays against the files written by the unmodified reference binary for the same inputs."""
    from tests.conftest import GOLDEN
    common.check_pipeline_vs_reference_files(gpu_ctx, small_workload, GOLDEN)


def test_cli_matches_golden_reference_files(small_workload):
    """`metamaps mapDirectly` + `classify` (C++ host over the CUDA library) vs the files of the unmodified reference."""
    import os
    from metamaps_b200 import build
    from tests import cli_common
    from tests.conftest import GOLDEN
    assert os.path.exists(build.HOST_BIN), "metamaps_b200/metamaps not built"
    got = cli_common.run_cli(build.HOST_BIN, small_workload["dir"], out="out_gpu")
    assert cli_common.compare_dirs(os.path.join(GOLDEN, "ref_small"), got) >= 8
    via = cli_common.run_cli_via_index(build.HOST_BIN, small_workload["dir"], out="out_gpu_ix")      # index + mapAgainstIndex
    cli_common.compare_mapping_files(got, via)
    assert cli_common.check_chunked_equals_direct(build.HOST_BIN, small_workload["dir"]) >= 3          # the --maxmemory chunk loop


def test_cli_config2_shaped_sample_matches_reference_files(tmp_path):
    """Config-2-shaped data from the reference itself: 408 Mbp of bench.py's DB recipe, 1500 reads, -m 2000 -w 16.  The
    reference's occurrence threshold is finite there (its own "ignore minimizers occurring >= 12 times" line, which the host
    must print too), so the over-frequent-hash skip (computeMap.hpp:314) and the L1 contig filter on a saturating 32-bit hash
    space are pinned by the unmodified reference's ten output files (tests/golden/ref_config2)."""
    import os
    from metamaps_b200 import build
    from tests import cli_common
    assert os.path.exists(build.HOST_BIN), "metamaps_b200/metamaps not built"
    identical, thr = cli_common.check_config2_golden(build.HOST_BIN, str(tmp_path))
    assert identical == 10, identical


@pytest.mark.parametrize("name", ["many_contigs", "long_read", "low_complexity"])
def test_fast_path_fallbacks(gpu_ctx, oracle, name):
    """Hashed contig bins, sketches too large for shared memory, 8-bit counter overflow: same results as the oracle."""
    contigs, reads, k, w, min_len = common.fallback_workloads()[name]
    common.check_map_vs_oracle(gpu_ctx, oracle, contigs, reads, k, w, 80.0, min_len, batches=1)


def test_multiple_l2_passes(oracle, small_workload, monkeypatch):
    """A tiny event budget forces the L2 stage through several passes; results must not change."""
    monkeypatch.setenv("MM_EV_BUDGET", "20000")
    ctx = capi.Context(0)
    contigs = [synth.codes_to_ascii(c) for c in small_workload["db"].contig_codes]
    reads = [synth.codes_to_ascii(r) for r in small_workload["reads"]]
    common.check_map_vs_oracle(ctx, oracle, contigs, reads, 16, 13)
    ctx.close()


@pytest.mark.parametrize("env", [{"MM_SWEEP_PRUNE": "0"}, {"MM_SWEEP_SEG": "100"}, {"MM_SWEEP_BAND": "64"}, {"MM_SWEEP_BAND": "128", "MM_SWEEP_RING": "4", "MM_SWEEP_SEG": "100"},
                                 {"MM_SWEEP_SEG": "64"}, {"MM_SWEEP_WIDE_FROM": "300"}, {"MM_SWEEP": "full"},
                                 {"MM_SWEEP": "global"}, {"MM_L1_FILTER": "legacy"}, {"MM_K3_GLOBAL": "1"}, {"MM_L1_SEGSORT": "0"}, {"MM_L1_CAND": "legacy"},
                                 {"MM_L1_FUSED": "0"}])
def test_kernel_variants(oracle, small_workload, monkeypatch, env):
    """Every selectable variant of K4/K5b (band width / ring depth of the banded sweep, the full-state shared-memory
    sweep, the global-memory sweep, the 8-byte L1 filter) gives the oracle's results."""
    for k_, v in env.items():
        monkeypatch.setenv(k_, v)
    ctx = capi.Context(0)
    contigs = [synth.codes_to_ascii(c) for c in small_workload["db"].contig_codes]
    reads = [synth.codes_to_ascii(r) for r in small_workload["reads"]]
    common.check_map_vs_oracle(ctx, oracle, contigs, reads, 16, 13)
    ctx.close()


def test_contig_shards_walked_in_one_process(gpu_ctx, small_workload):
    db = small_workload["db"]
    contigs = [synth.codes_to_ascii(c) for c in db.contig_codes]
    reads = [synth.codes_to_ascii(r) for r in small_workload["reads"]]
    taxa = sorted(set(db.contig_taxon)); tidx = {t: i for i, t in enumerate(taxa)}
    contig_taxon = np.array([tidx[t] for t in db.contig_taxon], np.int32)
    contig_len = np.array([len(c) for c in db.contig_codes], np.int64)
    common.check_shard_walk_equals_full(gpu_ctx, contigs, reads, 16, 13, contig_taxon, contig_len, len(taxa), cuts=[2, 5, 6])


def test_staged_reads(gpu_ctx, small_workload):
    contigs = [synth.codes_to_ascii(c) for c in small_workload["db"].contig_codes]
    common.check_staged_equals_direct(gpu_ctx, contigs, [synth.codes_to_ascii(r) for r in small_workload["reads"]], 16, 13)


def test_index_save_load(gpu_ctx, small_workload, tmp_path):
    contigs = [synth.codes_to_ascii(c) for c in small_workload["db"].contig_codes]
    common.check_index_save_load(gpu_ctx, contigs, [synth.codes_to_ascii(r) for r in small_workload["reads"]], 16, 13, str(tmp_path / "ix.0"))
    # same-hash-same-contig minimizers: the loaded index rebuilds its duplicate rank table (Index::build_dup_rank)
    from tests.test_kernel_logic_emu import _repetitive_workload
    rc, rr = _repetitive_workload()
    common.check_index_save_load(gpu_ctx, rc, rr, 16, 5, str(tmp_path / "ix.1"))


def test_map_at_scale_independent_paths_agree(oracle, monkeypatch):
    """Config-2-shaped workload at a size the oracle cannot finish (360 Mbp, 4 000 log-normal reads up to 40 kb, w = 16):
    the fast kernel chain (block-sort K3, 16-bit contig filter, window pruning, banded + segmented sweep) and the independent general chain
    (global sort, 8-byte filter, full-state sweep in global memory) must agree on every candidate; size-independent
    properties hold; a random subsample is checked against the oracle."""
    import torch
    import bench
    dev = torch.device("cuda", 0)
    wl = dict(n_species=60, n_strains=3, contig_len=2_000_000, div=0.01, n_reads=4000, mean_len=8000, sigma=0.5, min_read_len=2000, w=16, seed=17)
    asc, codes, offsets, _, _ = bench.gen_db(torch, dev, wl)
    r_asc, r_off = bench.gen_reads(torch, dev, wl, codes, 0)
    torch.cuda.synchronize()
    results = []
    for env in ({}, {"MM_SWEEP": "global", "MM_K3_GLOBAL": "1", "MM_L1_FILTER": "legacy"}):
        for k_, v in env.items():
            monkeypatch.setenv(k_, v)
        ctx = capi.Context(0)
        ix = capi.Index(ctx, 16, wl["w"])
        ix.add_dev(asc.data_ptr(), offsets); ix.finalize()
        res = capi.map_reads(ctx, ix, None, 80.0, wl["min_read_len"], dev_ptr=r_asc.data_ptr(), offsets=r_off, fetch_sketch=True)
        results.append(res)
        if env:
            assert res["stats"]["smem_swept"] == 0
        else:
            assert res["stats"]["smem_swept"] == len(res["shared"]) and res["stats"]["sweep_items"] > len(res["shared"])   # segments were cut
            assert 0 < res["stats"]["window_starts_swept"] < 0.5 * res["stats"]["window_starts"]                      # the prune pass dropped most window starts
        ix.close(); ctx.close()
    a, b = results
    assert len(a["shared"]) > 8000
    for key in ("s", "minimumHits", "cand_off", "seq", "start", "end", "pos", "shared", "votes", "accepted", "valid", "optStart", "optEnd", "q_off", "q_hash",
                "q_strand"):
        assert np.array_equal(a[key], b[key]), key
    # properties
    cand_read = np.repeat(np.arange(len(a["s"])), np.diff(a["cand_off"]))
    assert (a["shared"] <= a["s"][cand_read]).all() and (a["shared"] >= 0).all()
    assert (np.abs(a["votes"]) <= a["shared"]).all()
    assert (a["start"] <= a["end"]).all()
    same = cand_read[1:] == cand_read[:-1]
    key2 = a["seq"].astype(np.int64) * (1 << 32) + a["start"]
    assert (key2[1:][same] > key2[:-1][same]).all()                       # per read: (seqId, position) order, disjoint loci
    assert (np.diff(a["q_off"]) == a["s"]).all()
    for r in range(0, len(a["s"]), 97):                                   # sketches sorted and unique
        q = a["q_hash"][a["q_off"][r]:a["q_off"][r + 1]]
        assert (np.diff(q.astype(np.int64)) > 0).all()
    # oracle on a subsample (it needs the contigs the reads map to: use whole contigs of the first candidates)
    host_reads = r_asc.cpu().numpy()
    contig_ids = sorted(set(int(x) for x in a["seq"][:60]))[:6]
    L = wl["contig_len"]
    sub_contigs = [bytes(asc[c * L:(c + 1) * L].cpu().numpy()) for c in contig_ids]
    pick = [r for r in range(len(a["s"])) if a["cand_off"][r + 1] > a["cand_off"][r] and int(a["seq"][a["cand_off"][r]]) in contig_ids][:25]
    sub_reads = [bytes(host_reads[r_off[r]:r_off[r + 1]]) for r in pick]
    ctx = capi.Context(0)
    common.check_map_vs_oracle(ctx, oracle, sub_contigs, sub_reads, 16, wl["w"], 80.0, wl["min_read_len"], batches=1)
    ctx.close()


def test_cli_db_with_N_runs_lowercase_and_iupac(tmp_path):
    """The CUDA host binary against the unmodified reference on a DB with N runs, soft-masking and IUPAC codes."""
    from metamaps_b200 import build
    from tests import cli_common
    cli_common.check_db_with_N_runs_lowercase_and_iupac(build.HOST_BIN, tmp_path)


def test_api_errors(gpu_ctx, tmp_path):
    common.check_api_errors(gpu_ctx, tmp_path)


def test_staged_reads_pinned_zero_copy(gpu_ctx, small_workload, monkeypatch):
    """Pinned host memory through both staging modes -- DMA in pieces into a device slot (the default) and MM_STAGE=zerocopy (K0 packs
    straight from host memory over PCIe): same results as the direct call, including reads with non-ACGT bytes (the exception
    side list is settled at map time)."""
    import torch
    contigs = [synth.codes_to_ascii(c) for c in small_workload["db"].contig_codes]
    reads = [synth.codes_to_ascii(r) for r in small_workload["reads"][:120]]
    reads[3] = reads[3][:500] + b"N" * 40 + reads[3][540:]; reads[7] = reads[7].lower(); reads[11] = reads[11][:100] + b"RYK" + reads[11][103:]
    ix = common.build_index(gpu_ctx, contigs, 16, 13)
    direct = capi.map_reads(gpu_ctx, ix, reads, 80.0, 1000)
    off = np.zeros(len(reads) + 1, np.int64); off[1:] = np.cumsum([len(x) for x in reads])
    pinned = torch.empty(int(off[-1]) + 64, dtype=torch.uint8, pin_memory=True)
    pinned[:int(off[-1])] = torch.frombuffer(bytearray(b"".join(reads)), dtype=torch.uint8)
    for mode in (None, "zerocopy"):
        if mode:
            monkeypatch.setenv("MM_STAGE", mode)
        for slot in (0, 1):
            gpu_ctx.stage_reads(slot, pinned.data_ptr(), off)
        for slot in (0, 1):
            got = capi.map_reads(gpu_ctx, ix, None, 80.0, 1000, offsets=off, staged_slot=slot)
            for key in ("s", "cand_off", "seq", "pos", "shared", "votes", "accepted", "valid"):
                assert np.array_equal(got[key], direct[key]), (mode, slot, key)
            assert got["stats"]["exceptions"] == direct["stats"]["exceptions"] > 0


def test_pinned_result_buffers(gpu_ctx, small_workload):
    """capi.use_pinned_results: same pipeline results from the reused pinned buffers as from fresh arrays."""
    from metamaps_b200 import pipeline
    db = small_workload["db"]
    contigs = [synth.codes_to_ascii(c) for c in db.contig_codes]
    reads = [synth.codes_to_ascii(r) for r in small_workload["reads"]]
    taxa = sorted(set(db.contig_taxon)); tidx = {t: i for i, t in enumerate(taxa)}
    contig_taxon = np.array([tidx[t] for t in db.contig_taxon], np.int32)
    contig_len = np.array([len(c) for c in db.contig_codes], np.int64)
    ix = common.build_index(gpu_ctx, contigs, 16, 13)
    a = pipeline.map_and_classify(gpu_ctx, ix, reads=reads, contig_len=contig_len, contig_taxon=contig_taxon, n_taxa=len(taxa))
    a = {k_: (np.array(v) if isinstance(v, np.ndarray) else v) for k_, v in a.items()}
    fa, posta = a["em"]["f"].copy(), a["em"]["posterior"].copy()
    capi.use_pinned_results(True)
    try:
        for _ in range(2):          # second pass reuses the buffers
            b = pipeline.map_and_classify(gpu_ctx, ix, reads=reads, contig_len=contig_len, contig_taxon=contig_taxon, n_taxa=len(taxa))
            for key in common.MAPPING_KEYS:
                assert np.array_equal(a[key], b[key]), key
            # the EM taxon sums are atomics: their last bits depend on the order of the additions
            assert np.abs(fa - b["em"]["f"]).max() <= 1e-12 and np.abs(posta - b["em"]["posterior"]).max() <= 1e-12
    finally:
        capi.use_pinned_results(False)


def test_multi_batch_classify_table_and_top_mappings_filter(gpu_ctx, small_workload):
    common.check_multi_batch_classify(gpu_ctx, small_workload)


def test_cli_maxmemory_chunk_loop_matches_reference_fixture(tmp_path):
    """--maxmemory analogue on the GPU: three reference chunks with the reference's carried (non-reset) occurrence histogram,
    through mapDirectly's chunk loop and through a three-file persistent index; see cli_common.check_maxmemory_golden."""
    import os
    from metamaps_b200 import build
    from tests import cli_common
    assert os.path.exists(build.HOST_BIN), "metamaps_b200/metamaps not built"
    thr = cli_common.check_maxmemory_golden(build.HOST_BIN, str(tmp_path))
    assert [t.split(">= ")[1].split()[0] for t in thr] == ["4", "7", "23"]


def test_window_pruning_fuzz(monkeypatch):
    """The prune pass on the device (K5a's group counts, l2_prune_warp_kernel) against the sweep over every window start, on random
    repeat-rich references with k in {12, 14, 16}, w from 3 to 24 and reads with 0 - 15 % errors."""
    common.check_prune_fuzz(lambda: capi.Context(0), monkeypatch, 24)
