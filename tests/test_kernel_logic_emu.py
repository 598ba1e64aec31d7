"""Kernel logic on the CPU: the same functors the GPU runs, compiled with -DMM_HOST_EMU and executed in a
loop (tests/_emu/libmm_emu.so, test infrastructure only), checked against the oracle and the golden
fixtures.  This is what lets the build container (no GPU) catch logic errors; the parity tests proper are
in test_gpu_parity.py."""
import numpy as np
import pytest

from metamaps_b200 import capi, synth
from tests import common


def test_sketch_golden(emu_ctx, golden):
    common.check_sketch_vs_golden(emu_ctx, golden)


def test_sketch_edge_cases(emu_ctx, oracle):
    rng = np.random.default_rng(11)
    seqs = [b"", b"A", b"ACGTACGTACGTACG", b"ACGTACGTACGTACGT", b"N" * 300, b"ACGT" * 100 + b"N" * 64 + b"TTGACC" * 50,
            bytes(rng.choice(list(b"ACGT"), size=20001).astype(np.uint8)),
            bytes(rng.choice(list(b"acgtRYKM"), size=777).astype(np.uint8)), b"GATTACA" * 300]
    for (k, w) in ((16, 16), (16, 13), (16, 1), (15, 8), (9, 33), (4, 100), (16, 500)):
        common.check_sketch_vs_oracle(emu_ctx, oracle, seqs, k, w)


def test_sketch_long_monotone_window(emu_ctx, oracle):
    # w larger than the 32-entry register deque: exercises the global-memory replay of overflowing chunks
    rng = np.random.default_rng(5)
    seqs = [bytes(rng.choice(list(b"ACGT"), size=30000).astype(np.uint8))]
    common.check_sketch_vs_oracle(emu_ctx, oracle, seqs, 16, 2000)


@pytest.mark.parametrize("env", [{}, {"MM_SKETCH_CH": "32"}, {"MM_SKETCH_CH": "48"}, {"MM_SKETCH_CH": "128"},
                                 {"MM_SKETCH_BLOCKMIN": "0", "MM_SKETCH_CH": "32"}, {"MM_SKETCH_BAILCAP": "8", "MM_SKETCH_CH": "32"}])
def test_sketch_blockmin(emu_ctx, oracle, monkeypatch, env):
    """K1's block prefix / suffix minimum kernel and the chunks it hands back to the deque kernel, over chunk sizes that put
    the engineered runs on, before and after chunk boundaries; and the deque kernel alone over the same inputs."""
    for k_, v in env.items():
        monkeypatch.setenv(k_, v)
    seqs = common.blockmin_stress_seqs()
    for (k, w) in common.BLOCKMIN_PARAMS:
        common.check_sketch_vs_oracle(emu_ctx, oracle, seqs, k, w)


def test_index_golden(emu_ctx, golden, small_workload):
    common.check_index_vs_golden(emu_ctx, golden, small_workload)


def test_map_golden(emu_ctx, golden, small_workload):
    res = common.check_map_vs_golden(emu_ctx, golden, small_workload)
    assert res["summary"]["n_too_short"] == 4            # ref_small/ref.meta: ReadsTooShort 4
    assert res["summary"]["n_reads_mapped"] == 194       # ReadsMapped 194


def test_map_repetitive_reference(emu_ctx, oracle):
    """Tandem repeats and duplicated segments inside one contig: duplicate hashes inside L2 windows."""
    rng = np.random.default_rng(77)
    unit = rng.integers(0, 4, 700, dtype=np.uint8)
    parts = [rng.integers(0, 4, 5000, dtype=np.uint8), unit, unit, unit, rng.integers(0, 4, 3000, dtype=np.uint8), unit,
             rng.integers(0, 4, 4000, dtype=np.uint8)]
    c0 = np.concatenate(parts)
    c1 = np.concatenate([c0[2000:9000], rng.integers(0, 4, 2000, dtype=np.uint8), c0[2000:6000]])
    db = synth.SynthDB(["C0|kraken:taxid|1|x", "C1|kraken:taxid|2|x"], ["1", "2"], [c0, c1])
    _, reads, _ = synth.make_reads(db, 5, 60, 2500, err=0.06)
    contigs = [synth.codes_to_ascii(c) for c in db.contig_codes]
    res = common.check_map_vs_oracle(emu_ctx, oracle, contigs, [synth.codes_to_ascii(r) for r in reads], 16, 5, 80.0, 1000)
    assert res["summary"]["n_mappings"] > 0


def test_map_low_complexity_reads(emu_ctx, oracle):
    """Reads with repeated k-mers on both strands: the surviving duplicate must follow std::sort + std::unique."""
    rng = np.random.default_rng(9)
    genome = rng.integers(0, 4, 40000, dtype=np.uint8)
    pal = rng.integers(0, 4, 300, dtype=np.uint8)
    rc = (3 - pal[::-1]).astype(np.uint8)      # synth codes: A=0 C=1 G=2 T=3 -> complement = 3 - x
    genome[10000:10300] = pal; genome[10300:10600] = rc; genome[12000:12300] = pal
    db = synth.SynthDB(["C0|kraken:taxid|1|x"], ["1"], [genome])
    reads = [synth.codes_to_ascii(genome[9000:13500]), synth.codes_to_ascii(genome[9500:12800]),
             synth.codes_to_ascii((3 - genome[9000:13500][::-1]).astype(np.uint8))]
    common.check_map_vs_oracle(emu_ctx, oracle, [synth.codes_to_ascii(genome)], reads, 16, 4, 80.0, 1000, batches=1)


def test_map_ragged_and_empty(emu_ctx, oracle, small_workload):
    contigs = [synth.codes_to_ascii(c) for c in small_workload["db"].contig_codes]
    reads = [synth.codes_to_ascii(r) for r in small_workload["reads"][:12]]
    ragged = [b"", reads[0], b"ACGT", reads[1][:999], reads[2], b"N" * 1500, reads[3] + reads[4], reads[5][:1000]]
    common.check_map_vs_oracle(emu_ctx, oracle, contigs + [b"ACGT", b""], ragged, 16, 13)
    ix = common.build_index(emu_ctx, contigs, 16, 13)
    res = capi.map_reads(emu_ctx, ix, [], 80.0, 1000)
    assert res["summary"]["n_reads"] == 0 and res["summary"]["n_candidates"] == 0


def test_mapq(emu_ctx, oracle):
    common.check_mapq_vs_oracle(emu_ctx, oracle)


def test_em(emu_ctx, oracle):
    common.check_em_vs_oracle(emu_ctx, oracle, seed=5)
    common.check_em_vs_oracle(emu_ctx, oracle, seed=6, max_iter=50, nr=500, T=7, maxc=30)


def test_em_rejects_a_read_without_likelihood(emu_ctx):
    from metamaps_b200 import capi
    tax, mq, nloc, off, T = common.random_em_case(5, nr=50)
    mq = mq.copy(); mq[off[7]:off[8]] = 0.0
    with pytest.raises(capi.MMError):
        emu_ctx.em(tax, mq, nloc, off, T, 0)


def test_api_errors(emu_ctx, tmp_path):
    with pytest.raises(capi.MMError):
        emu_ctx.sketch([b"ACGT"], 17, 5)          # k > 16 (parseCmdArgs.hpp:62)
    with pytest.raises(capi.MMError):
        capi.Index(emu_ctx, 16, 0)
    common.check_api_errors(emu_ctx, tmp_path)


def test_pipeline_matches_reference_files(emu_ctx, small_workload):
    from tests.conftest import GOLDEN
    common.check_pipeline_vs_reference_files(emu_ctx, small_workload, GOLDEN)


@pytest.mark.parametrize("name", ["many_contigs", "low_complexity"])
def test_fallback_workloads(emu_ctx, oracle, name):
    contigs, reads, k, w, min_len = common.fallback_workloads()[name]
    common.check_map_vs_oracle(emu_ctx, oracle, contigs, reads, k, w, 80.0, min_len, batches=1)


def _repetitive_workload():
    rng = np.random.default_rng(77)
    unit = rng.integers(0, 4, 700, dtype=np.uint8)
    parts = [rng.integers(0, 4, 5000, dtype=np.uint8), unit, unit, unit, rng.integers(0, 4, 3000, dtype=np.uint8), unit,
             rng.integers(0, 4, 4000, dtype=np.uint8)]
    c0 = np.concatenate(parts)
    c1 = np.concatenate([c0[2000:9000], rng.integers(0, 4, 2000, dtype=np.uint8), c0[2000:6000]])
    db = synth.SynthDB(["C0|kraken:taxid|1|x", "C1|kraken:taxid|2|x"], ["1", "2"], [c0, c1])
    _, reads, _ = synth.make_reads(db, 5, 40, 2500, err=0.06)
    return [synth.codes_to_ascii(c) for c in db.contig_codes], [synth.codes_to_ascii(r) for r in reads]


@pytest.mark.parametrize("env", [{"MM_SWEEP_PRUNE": "0"}, {"MM_SWEEP_SEG": "100"}, {"MM_SWEEP_BAND": "64"}, {"MM_SWEEP_BAND": "128", "MM_SWEEP_SEG": "100"}, {"MM_SWEEP_SEG": "64"},
                                 {"MM_SWEEP": "global"}])
def test_sweep_variants(oracle, small_workload, monkeypatch, env):
    """K5b: a narrow band (the state is rebuilt from the window again and again) and the full-state sweep must both
    reproduce the reference's std::map sliding window, on plain and on repetitive references."""
    from metamaps_b200 import capi
    from tests.conftest import build_emu
    for k_, v in env.items():
        monkeypatch.setenv(k_, v)
    ctx = capi.Context(0, capi.load(build_emu()))
    contigs = [synth.codes_to_ascii(c) for c in small_workload["db"].contig_codes]
    reads = [synth.codes_to_ascii(r) for r in small_workload["reads"][:60]]
    common.check_map_vs_oracle(ctx, oracle, contigs, reads, 16, 13)
    rc, rr = _repetitive_workload()
    common.check_map_vs_oracle(ctx, oracle, rc, rr, 16, 5, 80.0, 1000)
    st = ctx.last_map_stats()
    if "MM_SWEEP" not in env:
        assert st["smem_swept"] > 0          # the banded path really ran
    ctx.close()


def test_contig_shards_walked_in_one_process(emu_ctx, small_workload):
    db = small_workload["db"]
    contigs = [synth.codes_to_ascii(c) for c in db.contig_codes]
    reads = [synth.codes_to_ascii(r) for r in small_workload["reads"][:80]]
    taxa = sorted(set(db.contig_taxon)); tidx = {t: i for i, t in enumerate(taxa)}
    contig_taxon = np.array([tidx[t] for t in db.contig_taxon], np.int32)
    contig_len = np.array([len(c) for c in db.contig_codes], np.int64)
    common.check_shard_walk_equals_full(emu_ctx, contigs, reads, 16, 13, contig_taxon, contig_len, len(taxa), cuts=[3, 7])


def test_staged_reads(emu_ctx, small_workload):
    contigs = [synth.codes_to_ascii(c) for c in small_workload["db"].contig_codes]
    common.check_staged_equals_direct(emu_ctx, contigs, [synth.codes_to_ascii(r) for r in small_workload["reads"][:60]], 16, 13)


def test_index_save_load(emu_ctx, small_workload, tmp_path):
    contigs = [synth.codes_to_ascii(c) for c in small_workload["db"].contig_codes]
    common.check_index_save_load(emu_ctx, contigs, [synth.codes_to_ascii(r) for r in small_workload["reads"][:40]], 16, 13, str(tmp_path / "ix.0"))
    # a reference with same-hash-same-contig minimizers: the loaded index rebuilds its duplicate rank table (Index::build_dup_rank)
    rc, rr = _repetitive_workload()
    common.check_index_save_load(emu_ctx, rc, rr, 16, 5, str(tmp_path / "ix.1"))


def test_multi_batch_classify_table_and_top_mappings_filter(emu_ctx, small_workload):
    common.check_multi_batch_classify(emu_ctx, small_workload)


def test_window_pruning_changes_nothing_but_the_work(oracle, monkeypatch):
    """K5b visits only the window starts whose upper bound reaches a lower bound of the optimum (l2_prune_core): same candidates,
    counts, positions and optimal windows as the full sweep and as the oracle, a fraction of the events."""
    import ctypes
    from tests.conftest import build_emu
    db = synth.make_db(7, 3, 3, 300_000, 0.01)
    _, reads, _ = synth.make_reads(db, 8, 120, 8000, lognormal_sigma=0.5, frac_random=0.05)
    contigs = [synth.codes_to_ascii(c) for c in db.contig_codes]; rd = [synth.codes_to_ascii(r) for r in reads]
    got = {}
    for flag in ("0", "1"):
        monkeypatch.setenv("MM_SWEEP_PRUNE", flag)
        lib = capi.load(build_emu()); ctx = capi.Context(0, lib)
        lib.mm_emu_sweep_iters.restype = ctypes.c_longlong
        lib.mm_emu_sweep_iters(1)
        res = common.check_map_vs_oracle(ctx, oracle, contigs, rd, 16, 16, 80.0, 2000, batches=1)
        got[flag] = (res, lib.mm_emu_sweep_iters(1))
    for key in ("seq", "pos", "shared", "votes", "accepted", "valid", "optStart", "optEnd"):
        assert np.array_equal(got["0"][0][key], got["1"][0][key]), key
    assert got["1"][1] < 0.4 * got["0"][1], (got["0"][1], got["1"][1])


def test_window_pruning_fuzz(monkeypatch):
    """Random references with tandem / dispersed repeats and near-identical contigs (duplicate hashes inside windows), k in {12, 14, 16},
    w from 3 to 24, reads with 0 - 15 % errors: the pruned sweep and the sweep over every window start agree on every output."""
    from tests.conftest import build_emu
    lib = capi.load(build_emu())
    common.check_prune_fuzz(lambda: capi.Context(0, lib), monkeypatch, 16)


def test_window_pruning_is_exact_whatever_its_ranks(monkeypatch):
    """The prune pass places its upper / lower ranks from an ESTIMATE of istar; its bounds hold only where their premises are checked to hold.
    Moving the ranks far off the estimate (test hook of the emulation build) must therefore change how much is swept, never a result."""
    import ctypes
    from tests.conftest import build_emu
    lib = capi.load(build_emu())
    lib.mm_emu_prune_shift.argtypes = [ctypes.c_int, ctypes.c_int]; lib.mm_emu_prune_shift.restype = None
    shifts = [(-400, 0), (0, 400), (-60, 60), (60, -60), (-25, 25), (300, -300), (-1000, -1000), (1000, 1000)]
    try:
        common.check_prune_fuzz(lambda: capi.Context(0, lib), monkeypatch, 16, before_pruned=lambda case: lib.mm_emu_prune_shift(*shifts[case % len(shifts)]))
    finally:
        lib.mm_emu_prune_shift(0, 0)

