"""World-size-2 (gloo, CPU) test of the multi-rank EM path: reads are partitioned over ranks, each rank runs mm_em_run
on its share and the per-round taxon sums are all-reduced through the host transport (mm_comm_set_allreduce).  The
result must equal the single-rank EM over all reads (the oracle)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from metamaps_b200 import capi
    from tests import common
    from tests.conftest import build_emu
    lib = capi.load(build_emu())
    ctx = capi.Context(0, lib)

    def allreduce(a):
        t = torch.from_numpy(a)
        dist.all_reduce(t)          # in place on the shared buffer
    ctx.set_allreduce(allreduce)
    tax, mq, nloc, off, T = common.random_em_case(31, nr=4000, T=60, maxc=15)
    # contiguous partition of the reads, like the bench shards read batches
    nr = len(off) - 1
    lo, hi = rank * nr // world, (rank + 1) * nr // world
    m0, m1 = off[lo], off[hi]
    res = ctx.em(tax[m0:m1], mq[m0:m1], nloc[m0:m1], off[lo:hi + 1] - m0, T)
    np.savez(os.path.join(tmp, f"rank{rank}.npz"), f=res["f"], post=res["posterior"], best=res["best"] + m0, iters=res["iters"], ll=res["ll"], lo=lo, hi=hi, m0=m0, m1=m1)
    dist.destroy_process_group()


def test_em_two_ranks_equals_single_rank(oracle, tmp_path):
    from tests import common
    from tests.conftest import build_emu
    build_emu()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    tax, mq, nloc, off, T = common.random_em_case(31, nr=4000, T=60, maxc=15)
    ref = oracle.em(tax, mq, nloc, off, T)
    parts = [np.load(os.path.join(str(tmp_path), f"rank{r}.npz")) for r in range(2)]
    for p in parts:
        assert int(p["iters"]) == ref["iters"]                        # same stopping round on every rank
        assert np.abs(p["f"] - ref["f"]).max() <= 1e-6                # identical global frequencies
        assert np.abs(p["ll"] - ref["ll"]).max() <= 1e-6 * np.abs(ref["ll"]).max()
        assert np.abs(p["post"] - ref["posterior"][int(p["m0"]):int(p["m1"])]).max() <= 1e-6
        assert np.array_equal(p["best"], ref["best"][int(p["lo"]):int(p["hi"])])
    assert np.array_equal(parts[0]["f"], parts[1]["f"])


# ---------------------------------------------------------------------------------------------- contig-range shards
def _shard_worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    os.environ["MM_SYNC_STEP"] = "50000"            # several hash ranges even on this small reference
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from metamaps_b200 import capi, pipeline
    from tests import common
    from tests.conftest import build_emu
    lib = capi.load(build_emu())
    ctx = capi.Context(0, lib)

    def allreduce(a):
        dist.all_reduce(torch.from_numpy(a))
    ctx.set_allreduce(allreduce); ctx.set_rank(world, rank)
    db, contigs, reads, w, contig_taxon, contig_len, T = common.sharded_case()
    per = len(contigs) // world
    c0, c1 = rank * per, (rank + 1) * per if rank + 1 < world else len(contigs)
    ix = capi.Index(ctx, 16, w); ix.set_shard(c0, keep_counts=True); ix.add(contigs[c0:c1]); ix.finalize()
    local_thr = ix.stats()["freq_threshold"]
    thr, uniq = ix.sync_threshold()

    n = len(reads); lo, hi = rank * n // world, (rank + 1) * n // world
    res = pipeline.map_and_classify_sharded(ctx, [ix], reads=reads, contig_len=contig_len, contig_taxon=contig_taxon, n_taxa=T,
                                            read_range=(lo, hi))
    np.savez(os.path.join(tmp, f"shard{rank}.npz"), thr=thr, uniq=uniq, local_thr=local_thr, lo=lo, hi=hi, f=res["em"]["f"], iters=res["em"]["iters"],
             post=res["em"]["posterior"], **{k_: res[k_] for k_ in common.MAPPING_KEYS})
    # sketch-once variant: the rank holds only ITS (unequal) block of the reads; sketches are all-gathered (mm_map_batch_sharded_dev)
    cut = 25
    mine = reads[:cut] if rank == 0 else reads[cut:]
    data = np.frombuffer(b"".join(mine), np.uint8).copy()
    off = np.zeros(len(mine) + 1, np.int64); off[1:] = np.cumsum([len(x) for x in mine])
    res2 = pipeline.map_and_classify_sharded(ctx, [ix], contig_len=contig_len, contig_taxon=contig_taxon, n_taxa=T, my_block=(data.ctypes.data, off))
    assert res2["read_range"] == ((0, cut) if rank == 0 else (cut, n)), res2["read_range"]
    assert res2["summary"]["n_reads"] == n
    np.savez(os.path.join(tmp, f"shardB{rank}.npz"), lo=res2["read_range"][0], hi=res2["read_range"][1], f=res2["em"]["f"], iters=res2["em"]["iters"],
             post=res2["em"]["posterior"], **{k_: res2[k_] for k_ in common.MAPPING_KEYS})
    dist.destroy_process_group()


def test_contig_shards_two_ranks_equal_unsharded_reference(tmp_path):
    """World size 2 (gloo): each rank indexes half of the contigs, the occurrence threshold of the WHOLE reference is
    agreed through mm_index_sync_threshold, every rank maps all reads against its shard, mappings are exchanged, each
    rank finalises its block of reads and the EM sums are all-reduced.  Everything must equal the one-index run."""
    from metamaps_b200 import capi, pipeline
    from tests import common
    from tests.conftest import build_emu
    lib = capi.load(build_emu())
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_shard_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    ctx = capi.Context(0, lib)
    db, contigs, reads, w, contig_taxon, contig_len, T = common.sharded_case()
    full = common.build_index(ctx, contigs, 16, w)
    st = full.stats()
    assert st["freq_threshold"] == 70                    # finite: over-frequent hashes exist ...
    ref = pipeline.map_and_classify(ctx, full, reads=reads, contig_len=contig_len, contig_taxon=contig_taxon, n_taxa=T)
    parts = [np.load(os.path.join(str(tmp_path), f"shard{r}.npz")) for r in range(2)]
    for p in parts:
        assert int(p["thr"]) == st["freq_threshold"] and int(p["uniq"]) == st["n_unique"]
        assert int(p["local_thr"]) != st["freq_threshold"]          # ... which no shard can see on its own
        sel = (ref["read"] >= int(p["lo"])) & (ref["read"] < int(p["hi"]))
        for key in common.MAPPING_KEYS:
            assert np.array_equal(p[key], ref[key][sel]), key          # coordinates, counts, identities, mapq: identical
        assert int(p["iters"]) == ref["em"]["iters"]
        assert np.abs(p["f"] - ref["em"]["f"]).max() <= 1e-6
        assert np.abs(p["post"] - ref["em"]["posterior"][sel]).max() <= 1e-6
    assert np.array_equal(parts[0]["f"], parts[1]["f"])
    assert sum(len(p["read"]) for p in parts) == len(ref["read"])
    partsB = [np.load(os.path.join(str(tmp_path), f"shardB{r}.npz")) for r in range(2)]
    for p in partsB:                                                    # sketch once + all-gather of the sketches: the same results
        sel = (ref["read"] >= int(p["lo"])) & (ref["read"] < int(p["hi"]))
        for key in common.MAPPING_KEYS:
            assert np.array_equal(p[key], ref[key][sel]), key
        assert int(p["iters"]) == ref["em"]["iters"] and np.abs(p["f"] - ref["em"]["f"]).max() <= 1e-6
        assert np.abs(p["post"] - ref["em"]["posterior"][sel]).max() <= 1e-6
    assert sum(len(p["read"]) for p in partsB) == len(ref["read"])


# ---------------------------------------------------------------------------------------------- streamed, interleaved chunks
def _stream_worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from metamaps_b200 import capi, pipeline
    from tests.conftest import build_emu
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    lib = capi.load(build_emu())
    ctx = capi.Context(0, lib)

    def allreduce(a):
        dist.all_reduce(torch.from_numpy(a))
    ctx.set_allreduce(allreduce); ctx.set_rank(world, rank)
    contigs, reads, cuts, contig_taxon, contig_len, T = _stream_case()
    bounds = np.concatenate([[0], np.cumsum(cuts)])

    def build_chunk(c):
        ix = capi.Index(ctx, 16, 10); ix.set_shard(int(bounds[c]), keep_counts=False); ix.add(contigs[bounds[c]:bounds[c + 1]]); ix.finalize()
        return ix

    def all_gather(obj):
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out
    n = len(reads); lo, hi = rank * n // world, (rank + 1) * n // world
    res = pipeline.map_and_classify_streamed(ctx, build_chunk, list(range(rank, len(cuts), world)), len(cuts), all_gather, reads=reads, contig_len=contig_len,
                                             contig_taxon=contig_taxon, n_taxa=T, read_range=(lo, hi))
    from tests import common
    np.savez(os.path.join(tmp, f"stream{rank}.npz"), lo=lo, hi=hi, f=res["em"]["f"], iters=res["em"]["iters"], thr=np.array(sorted(res["thresholds"].items())),
             **{k_: res[k_] for k_ in common.MAPPING_KEYS})
    dist.destroy_process_group()


def _stream_case():
    """Five chunks of a small reference with planted repeats, so that the carried (non-reset) histogram changes the thresholds."""
    rng = np.random.default_rng(17)
    from metamaps_b200 import synth
    n_contigs, L = 10, 600_000
    contigs = [rng.integers(0, 4, L, dtype=np.uint8) for _ in range(n_contigs)]
    for copies in (60, 45, 30):
        u = rng.integers(0, 4, 25, dtype=np.uint8)
        for _ in range(copies):
            ci = int(rng.integers(0, n_contigs)); pos = int(rng.integers(0, L - 50)); contigs[ci][pos:pos + 25] = u
    db = synth.SynthDB([f"C{i}|kraken:taxid|{i + 1}|x" for i in range(n_contigs)], [str(i + 1) for i in range(n_contigs)], contigs)
    _, reads, _ = synth.make_reads(db, 18, 50, 2500, err=0.08)
    return ([synth.codes_to_ascii(c) for c in contigs], [synth.codes_to_ascii(r) for r in reads], [2, 2, 2, 2, 2],
            np.arange(n_contigs, dtype=np.int32), np.full(n_contigs, L, np.int64), n_contigs)


def test_streamed_interleaved_chunks_two_ranks_equal_the_one_process_chunk_walk(tmp_path):
    """Config 5's driver logic at world size 2 (gloo): rank g owns chunks g, g+2, ...; after every round of builds the ranks
    exchange the chunks' own occurrence histograms and settle the reference's carried thresholds; all reads are mapped against
    every chunk; the tables are merged by (read, contig).  Must equal ONE process walking the five chunks in order."""
    from metamaps_b200 import capi, pipeline
    from tests import common
    from tests.conftest import build_emu
    lib = capi.load(build_emu())
    port = 33500 + (os.getpid() % 2000)
    mp.spawn(_stream_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    ctx = capi.Context(0, lib)
    contigs, reads, cuts, contig_taxon, contig_len, T = _stream_case()
    bounds = np.concatenate([[0], np.cumsum(cuts)])

    def build_chunk(c):
        ix = capi.Index(ctx, 16, 10); ix.set_shard(int(bounds[c]), keep_counts=False); ix.add(contigs[bounds[c]:bounds[c + 1]]); ix.finalize()
        return ix
    ref = pipeline.map_and_classify_streamed(ctx, build_chunk, list(range(len(cuts))), len(cuts), lambda o: [o], reads=reads, contig_len=contig_len,
                                             contig_taxon=contig_taxon, n_taxa=T)
    thr = sorted(ref["thresholds"].items())
    assert len(set(t for _, t in thr)) > 1, thr            # the carried histogram really changes the thresholds along the chain
    parts = [np.load(os.path.join(str(tmp_path), f"stream{r}.npz")) for r in range(2)]
    seen = {}
    for p in parts:
        seen.update({int(a): int(b) for a, b in p["thr"]})
        sel = (ref["read"] >= int(p["lo"])) & (ref["read"] < int(p["hi"]))
        for key in common.MAPPING_KEYS:
            assert np.array_equal(p[key], ref[key][sel]), key
        assert int(p["iters"]) == ref["em"]["iters"] and np.abs(p["f"] - ref["em"]["f"]).max() <= 1e-6
    assert sorted(seen.items()) == thr
