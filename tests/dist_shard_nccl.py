"""Run under torchrun (N >= 2 GPUs, NCCL): contig-range shards, one per GPU, against the one-index run on the same GPU.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/dist_shard_nccl.py
Exit code 0 = every rank's block of reads matches the unsharded result (tests/test_multi_gpu.py wraps this)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from metamaps_b200 import capi, pipeline
    from tests import common
    world = int(os.environ["WORLD_SIZE"]); rank = int(os.environ["RANK"]); local = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("MM_SYNC_STEP", "50000")
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = capi.Context(local)
    uid = [ctx.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctx.comm_init(world, rank, uid[0])
    db, contigs, reads, w, contig_taxon, contig_len, T = common.sharded_case()
    per = len(contigs) // world
    c0, c1 = rank * per, (rank + 1) * per if rank + 1 < world else len(contigs)
    ix = capi.Index(ctx, 16, w); ix.set_shard(c0, keep_counts=True); ix.add(contigs[c0:c1]); ix.finalize()
    thr, uniq = ix.sync_threshold()

    n = len(reads); lo, hi = rank * n // world, (rank + 1) * n // world
    res = pipeline.map_and_classify_sharded(ctx, [ix], reads=reads, contig_len=contig_len, contig_taxon=contig_taxon, n_taxa=T,
                                            read_range=(lo, hi))
    # the one-index run, on a second context without a communicator (its EM must not wait for the other ranks)
    ctx1 = capi.Context(local)
    full = common.build_index(ctx1, contigs, 16, w)
    st = full.stats()
    ref = pipeline.map_and_classify(ctx1, full, reads=reads, contig_len=contig_len, contig_taxon=contig_taxon, n_taxa=T)
    assert thr == st["freq_threshold"] == 70 and uniq == st["n_unique"], (thr, uniq, st)
    sel = (ref["read"] >= lo) & (ref["read"] < hi)
    for key in common.MAPPING_KEYS:
        assert np.array_equal(res[key], ref[key][sel]), key
    assert res["em"]["iters"] == ref["em"]["iters"]
    assert np.abs(res["em"]["f"] - ref["em"]["f"]).max() <= 1e-6
    assert np.abs(res["em"]["posterior"] - ref["em"]["posterior"][sel]).max() <= 1e-6
    # sketch once: the rank holds only its (unequal) block of the reads on the device; the sketches are all-gathered over NCCL
    cut = 25
    mine = reads[:cut] if rank == 0 else reads[cut:]
    blk = torch.frombuffer(bytearray(b"".join(mine)), dtype=torch.uint8).cuda()
    off = np.zeros(len(mine) + 1, np.int64); off[1:] = np.cumsum([len(x) for x in mine])
    res2 = pipeline.map_and_classify_sharded(ctx, [ix], contig_len=contig_len, contig_taxon=contig_taxon, n_taxa=T, my_block=(blk.data_ptr(), off))
    lo2, hi2 = res2["read_range"]
    assert (lo2, hi2) == ((0, cut) if rank == 0 else (cut, n)), (lo2, hi2)
    sel2 = (ref["read"] >= lo2) & (ref["read"] < hi2)
    for key in common.MAPPING_KEYS:
        assert np.array_equal(res2[key], ref[key][sel2]), ("sketch-once", key)
    assert res2["em"]["iters"] == ref["em"]["iters"] and np.abs(res2["em"]["f"] - ref["em"]["f"]).max() <= 1e-6
    print(f"rank {rank}: contig shards [{c0},{c1}) ok, global threshold {thr}, {int(sel.sum())} mappings of reads [{lo},{hi})", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
