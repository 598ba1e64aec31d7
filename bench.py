#!/usr/bin/env python
"""bench.py -- long-read Mbp/s, map + EM-classify (BASELINE.json metric), on N B200s of one node.

A "step" is one pass of the hot path (K1 sketch -> K3 read sketch -> K4 L1 -> K5 L2 -> K6 mapq -> K7/K8 EM)
over one batch of synthetic reads against a GPU-resident index of a synthetic DB.

  value   reads already resident in HBM as ASCII when the timed region starts (mm_map_batch_dev)
  e2e     the same metric through the host-buffer C-ABI calls (mm_map_batch on pinned host reads), H2D of the
          reads and D2H of every result array inside the timed region
N > 1 (torchrun): the index is replicated, reads are sharded (weak scaling: every rank maps its own batch),
the only collective is the per-round NCCL all-reduce of the EM taxon sums (mm_comm_*).

--impl reference: the unmodified reference CLI built with the Boost shim (oracle/_ref/metamaps), all host
threads, on a bounded sample of the same workload (see cpu_baseline.sample in the output line).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout carries the one JSON line: the image's NCCL_DEBUG=VERSION makes NCCL print its banner there
if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"

WORKLOADS = {
    # BASELINE.json configs[1]: 100k reads (mean 8 kb, log-normal sigma 0.5) vs 1000 species x 3 strains x 4 Mbp = 12 Gbp
    "config2": dict(n_species=1000, n_strains=3, contig_len=4_000_000, div=0.01, n_reads=100_000, mean_len=8000,
                    sigma=0.5, min_read_len=2000, w=16, seed=11),
    # a scaled-down copy for quick checks (same shape, 1.2 Gbp / 10k reads)
    "config2-small": dict(n_species=100, n_strains=3, contig_len=4_000_000, div=0.01, n_reads=10_000, mean_len=8000,
                          sigma=0.5, min_read_len=2000, w=16, seed=11),
    "tiny": dict(n_species=8, n_strains=3, contig_len=500_000, div=0.01, n_reads=2_000, mean_len=8000,
                 sigma=0.5, min_read_len=2000, w=16, seed=11),
}
K = 16
PI = 80.0


# ------------------------------------------------------------------------------------------ synthetic data
def gen_db(torch, dev, wl):
    """ASCII DB on the device: n_species ancestors, each strain = ancestor with `div` substitutions."""
    g = torch.Generator(device=dev); g.manual_seed(wl["seed"])
    n_contigs = wl["n_species"] * wl["n_strains"]; L = wl["contig_len"]
    asc = torch.empty(n_contigs * L, dtype=torch.uint8, device=dev)
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
    codes = torch.empty(n_contigs * L, dtype=torch.uint8, device=dev)
    ci = 0
    for s in range(wl["n_species"]):
        anc = torch.randint(0, 4, (L,), dtype=torch.uint8, device=dev, generator=g)
        for t in range(wl["n_strains"]):
            mut = torch.rand(L, device=dev, generator=g) < wl["div"]
            sub = torch.randint(1, 4, (L,), dtype=torch.uint8, device=dev, generator=g)
            c = torch.where(mut, (anc + sub) & 3, anc)
            codes[ci * L:(ci + 1) * L] = c
            asc[ci * L:(ci + 1) * L] = lut[c.long()]
            ci += 1
    offsets = np.arange(n_contigs + 1, dtype=np.int64) * L
    contig_taxon = np.arange(n_contigs, dtype=np.int32)          # one taxon (strain) per contig
    contig_len = np.full(n_contigs, L, np.int64)
    return asc, codes, offsets, contig_taxon, contig_len


def gen_reads(torch, dev, wl, codes, rank, err=0.12):
    """Reads sampled from the DB with 12 % sub/ins/del errors (simulate.pl:57), random strand; ASCII on the device."""
    rng = np.random.Generator(np.random.PCG64(wl["seed"] * 1000 + rank))
    n = wl["n_reads"]; L = wl["contig_len"]; n_contigs = wl["n_species"] * wl["n_strains"]
    mu = np.log(wl["mean_len"]) - 0.5 * wl["sigma"] ** 2
    lens = np.clip(rng.lognormal(mu, wl["sigma"], n), 1200, 40000).astype(np.int64)
    contig = rng.integers(0, n_contigs, n); start = (rng.random(n) * (L - lens)).astype(np.int64)
    rev = rng.random(n) < 0.5
    g = torch.Generator(device=dev); g.manual_seed(wl["seed"] * 7919 + rank)
    src_off = np.zeros(n + 1, np.int64); src_off[1:] = np.cumsum(lens)
    tot = int(src_off[-1])
    t_len = torch.from_numpy(lens).to(dev); t_off = torch.from_numpy(src_off[:-1]).to(dev)
    t_base = torch.from_numpy(contig * L + start).to(dev); t_rev = torch.from_numpy(rev).to(dev)
    rid = torch.repeat_interleave(torch.arange(n, device=dev), t_len)
    j = torch.arange(tot, device=dev) - t_off[rid]
    pos = torch.where(t_rev[rid], t_base[rid] + (t_len[rid] - 1 - j), t_base[rid] + j)
    b = codes[pos]
    b = torch.where(t_rev[rid], 3 - b, b)            # codes here: A0 C1 G2 T3 -> complement = 3 - x
    u = torch.rand(tot, device=dev, generator=g)
    e3 = err / 3
    sub = u < e3; ins = (u >= e3) & (u < 2 * e3); dele = (u >= 2 * e3) & (u < 3 * e3)
    b = torch.where(sub, (b + torch.randint(1, 4, (tot,), dtype=torch.uint8, device=dev, generator=g)) & 3, b)
    out_len = torch.ones(tot, dtype=torch.int64, device=dev); out_len[ins] = 2; out_len[dele] = 0
    opos = torch.cumsum(out_len, 0) - out_len
    total_out = int(out_len.sum().item())
    out = torch.empty(total_out, dtype=torch.uint8, device=dev)
    keep = ~dele
    out[opos[keep]] = b[keep]
    out[opos[ins] + 1] = torch.randint(0, 4, (int(ins.sum().item()),), dtype=torch.uint8, device=dev, generator=g)
    # per-read output lengths
    csum = torch.cat([torch.zeros(1, dtype=torch.int64, device=dev), torch.cumsum(out_len, 0)])
    roff = csum[torch.from_numpy(src_off).to(dev)].cpu().numpy().astype(np.int64)
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
    asc = lut[out.long()]
    return asc, roff


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index; self.stop_flag = False; self.sm = []; self.reasons = set(); self.max_sm = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml; self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if not self.nv:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self.stop_flag:
            try:
                self.sm.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.1)

    def result(self):
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_sm, "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------------------------------ reference arm
def reference_sample(wl, n_contigs, n_reads, outdir):
    """A bounded sample of the workload for the CPU reference: the first n_contigs contigs of the same DB recipe
    and n_reads reads drawn from them with the same length/error model (numpy, seeded)."""
    from metamaps_b200 import synth
    sp = max(1, n_contigs // wl["n_strains"])
    db = synth.make_db(wl["seed"], sp, wl["n_strains"], wl["contig_len"], wl["div"])
    fa = synth.write_db(db, os.path.join(outdir, "db"))
    names, reads, _ = synth.make_reads(db, wl["seed"] + 1, n_reads, wl["mean_len"], lognormal_sigma=wl["sigma"])
    fq = os.path.join(outdir, "reads.fq")
    synth.write_fastq(fq, names, reads)
    bases = int(sum(len(r) for r in reads if len(r) >= wl["min_read_len"]))
    return fa, fq, bases, len(db.contig_codes)


def run_reference_once(binary, d, wl, threads):
    """mapDirectly + classify with the reference CLI.  Returns (seconds spent mapping + classifying, index seconds).
    The reference builds its index inside mapDirectly; its own log line gives the mapping time."""
    import re
    out = os.path.join(d, "out"); os.makedirs(out, exist_ok=True)
    t0 = time.time()
    p = subprocess.run([binary, "mapDirectly", "--all", "-r", "db/DB.fa", "-q", "reads.fq", "-o", "out/ref", "-m", str(wl["min_read_len"]),
                        "-w", str(wl["w"]), "-t", str(threads)], cwd=d, capture_output=True, text=True)
    t_map_total = time.time() - t0
    if p.returncode != 0:
        raise RuntimeError("reference mapDirectly failed: " + p.stderr[-500:])
    m = re.search(r"Time spent mapping the query : ([0-9.eE+-]+) sec", p.stdout)
    t_map = float(m.group(1)) if m else t_map_total
    t1 = time.time()
    p = subprocess.run([binary, "classify", "--DB", "db", "--mappings", "out/ref", "-t", str(threads)], cwd=d, capture_output=True, text=True)
    t_cls = time.time() - t1
    if p.returncode != 0:
        raise RuntimeError("reference classify failed: " + p.stderr[-500:])
    return t_map + t_cls, t_map_total - t_map


def reference_arm(args, wl):
    from oracle import pyoracle
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    base = {"impl": "reference", "metric": "long_read_Mbp_per_s_map_plus_EM_classify", "unit": "Mbp/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int64/f64", "data": "synthetic"}
    if not os.path.exists(pyoracle.REF_BIN):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/metamaps missing (build it where /root/reference exists)"}))
        return
    d = tempfile.mkdtemp(prefix="mmref_")
    n_contigs, n_reads = args.ref_contigs, args.ref_reads
    fa, fq, bases, nc = reference_sample(wl, n_contigs, n_reads, d)
    times = []
    for i in range(args.warmup + args.steps):
        t, t_index = run_reference_once(pyoracle.REF_BIN, d, wl, threads)
        if i >= args.warmup:
            times.append(t)
    sec = float(np.mean(times))
    val = bases / 1e6 / sec
    sample = (f"{n_reads} reads (mean {wl['mean_len']} b, log-normal) vs the first {nc} contigs ({nc * wl['contig_len'] / 1e6:.0f} Mbp) of the "
              f"{args.workload} DB recipe; reference CLI mapDirectly (index build excluded, its own 'Time spent mapping' line) + classify, -t {threads}")
    base.update({"value": val, "ms_per_step": sec * 1e3, "config": {"workload": args.workload, "sample": sample},
                 "cpu_baseline": {"value": val, "unit": "Mbp/s", "cores": threads, "kind": "reference", "sample": sample},
                 "e2e": {"value": val, "unit": "Mbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    print(json.dumps(base))


# ------------------------------------------------------------------------------------------ main arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="config2", choices=sorted(WORKLOADS))
    ap.add_argument("--reads", type=int, default=0, help="override reads per GPU")
    ap.add_argument("--ref-contigs", type=int, default=12, help="reference arm: DB sample size in contigs")
    ap.add_argument("--ref-reads", type=int, default=1500, help="reference arm: reads in the sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--in-flight", type=int, default=1,
                    help="batches in flight per GPU (host threads, one context each, sharing the index); measured on config 2: 2 in flight = "
                         "+3 % value, -4 % e2e with the session-2 kernels and -7 % / -15 % with the session-3 ones, so the default stays 1")
    ap.add_argument("--shard", default="reads", choices=["reads", "contigs"],
                    help="N > 1: 'reads' = index replicated, every rank maps its own reads (no mapping exchange); 'contigs' = the "
                         "index is split into contig ranges, every rank maps ALL reads against its shard, mappings are exchanged")
    ap.add_argument("--e2e-breakdown", action="store_true",
                    help="after the timed runs: three more e2e steps with the per-stage CUDA-event times and the wall-clock split of each "
                         "printed to stderr next to those of a device-resident step (what staging the next batch costs the current one)")
    ap.add_argument("--profile-step", action="store_true",
                    help="after the warm-up run ONE step between cudaProfilerStart/Stop and exit (for `ncu --profile-from-start off`)")
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.reads:
        wl["n_reads"] = args.reads
    if args.impl == "reference":
        reference_arm(args, wl)
        return

    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:       # the host helpers (identities, nLoc) are OpenMP loops: share the cores between the ranks of the node
        os.environ.setdefault("OMP_NUM_THREADS", str(max(1, (os.cpu_count() or 1) // world)))
        os.environ.setdefault("OMP_WAIT_POLICY", "PASSIVE")       # idle OpenMP workers must not spin on cores the other ranks need
    import torch
    import torch.distributed as dist
    from metamaps_b200 import capi, pipeline
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = capi.Context(local)                       # raises if the CUDA library / device is missing: no fallback
    capi.use_pinned_results(True)                   # like a streaming host: result arrays live in reused pinned buffers
    if world > 1:
        uid = [ctx.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(world, rank, uid[0])

    t_setup = time.time()
    asc, codes, offsets, contig_taxon, contig_len = gen_db(torch, dev, wl)
    torch.cuda.synchronize()
    ix = capi.Index(ctx, K, wl["w"])
    n_contigs = len(offsets) - 1
    step_c = max(1, (256_000_000 // wl["contig_len"]))
    t_ix = time.time()
    by_contigs = args.shard == "contigs" and world > 1
    own0, own1 = (rank * n_contigs // world, (rank + 1) * n_contigs // world) if by_contigs else (0, n_contigs)
    if by_contigs:
        ix.set_shard(own0, keep_counts=True)
    for c0 in range(own0, own1, step_c):
        c1 = min(own1, c0 + step_c)
        ix.add_dev(asc.data_ptr(), offsets[c0:c1 + 1])
    del asc
    ix.finalize()
    if by_contigs:
        ix.sync_threshold()                        # collective: occurrence threshold of the whole reference
    index_s = time.time() - t_ix
    istats = ix.stats()
    if by_contigs:                                 # every rank needs every read: blocks 0..N-1 of the read recipe
        blocks = [gen_reads(torch, dev, wl, codes, b) for b in range(world)]
        r_asc = torch.cat([b[0] for b in blocks])
        r_off = np.concatenate([[0]] + [b[1][1:] + sum(int(x[1][-1]) for x in blocks[:i]) for i, b in enumerate(blocks)]).astype(np.int64)
        blk = np.cumsum([0] + [len(b[1]) - 1 for b in blocks])
        my_reads = (int(blk[rank]), int(blk[rank + 1]))
        del blocks
    else:
        r_asc, r_off = gen_reads(torch, dev, wl, codes, rank)
    del codes
    torch.cuda.empty_cache()
    n_taxa = int(contig_taxon.max()) + 1
    read_len = np.diff(r_off).astype(np.int32)
    # pinned host copy of the reads for the e2e leg
    r_host = torch.empty(r_asc.numel(), dtype=torch.uint8, pin_memory=True)
    r_host.copy_(r_asc); torch.cuda.synchronize()
    setup_s = time.time() - t_setup

    MKEYS = (("read", np.int32), ("seq", np.int32), ("pos", np.int32), ("shared", np.int32), ("sketch", np.int32), ("strand", np.int32),
             ("identity", np.float32), ("identity_parsed", np.float64))

    def exchange(parts):
        """all-gather of this rank's accepted mappings (one shard per rank): sizes first, then one padded byte tensor over NCCL."""
        m = parts[0]
        n = len(m["read"])
        sizes = torch.zeros(world, dtype=torch.int64, device=dev); sizes[rank] = n
        dist.all_reduce(sizes)
        sizes = sizes.cpu().numpy(); cap = int(sizes.max())
        rec = sum(np.dtype(t).itemsize for _, t in MKEYS)
        buf = np.zeros(cap * rec, np.uint8); o = 0
        for k_, t in MKEYS:
            b = np.ascontiguousarray(m[k_], t).view(np.uint8); buf[o:o + b.size] = b; o += cap * np.dtype(t).itemsize
        send = torch.from_numpy(buf).to(dev)
        recv = torch.empty(world * cap * rec, dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(recv, send)
        allb = recv.cpu().numpy().reshape(world, cap * rec)
        out = []
        for r in range(world):
            d = {}; o = 0
            for k_, t in MKEYS:
                isz = np.dtype(t).itemsize
                d[k_] = allb[r, o:o + int(sizes[r]) * isz].view(t).copy(); o += cap * isz
            out.append([d])
        return out

    def step_dev(stats=None):
        if by_contigs:
            return pipeline.map_and_classify_sharded(ctx, [ix], dev_ptr=r_asc.data_ptr(), offsets=r_off, read_len=read_len, contig_len=contig_len,
                                                     contig_taxon=contig_taxon, n_taxa=n_taxa, perc_identity=PI, min_read_len=wl["min_read_len"],
                                                     exchange=exchange, read_range=my_reads, stats=stats)
        return pipeline.map_and_classify(ctx, ix, dev_ptr=r_asc.data_ptr(), offsets=r_off, read_len=read_len, contig_len=contig_len,
                                         contig_taxon=contig_taxon, n_taxa=n_taxa, perc_identity=PI, min_read_len=wl["min_read_len"], stats=stats)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        out = step_dev()
    if args.profile_step:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        step_dev()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return

    # Two batches in flight: two host threads, each with its own context (own streams and scratch) mapping against the
    # shared read-only index, so that one batch's kernels fill the other's host-synchronisation bubbles.  The classify stage
    # (mapq, nLoc, EM) runs on the main context in step order (a ticket), which keeps the EM all-reduces of the ranks aligned.
    n_flight = 1 if by_contigs else max(1, args.in_flight)
    wctx = [capi.Context(local) for _ in range(n_flight)] if n_flight > 1 else [ctx]
    cond = threading.Condition(); turn = [0]

    def pipelined_step(i, w, staged, stats=None):
        t0_ = time.perf_counter()
        res = capi.map_reads(wctx[w], ix, None, PI, wl["min_read_len"], dev_ptr=None if staged is not None else r_asc.data_ptr(), offsets=r_off,
                             fetch=False, staged_slot=staged)
        t1_ = time.perf_counter()
        m = capi.fetch_mappings(wctx[w], res["summary"]["n_mappings"])
        t2_ = time.perf_counter()
        with cond:
            cond.wait_for(lambda: turn[0] == i)
        wall_ = {"map_call": (t1_ - t0_) * 1e3, "fetch_mappings": (t2_ - t1_) * 1e3}
        o = pipeline.classify_mappings(ctx, m, res["_n"], read_len, k=K, contig_len=contig_len, contig_taxon=contig_taxon, n_taxa=n_taxa, wall=wall_)
        with cond:
            turn[0] += 1; cond.notify_all()
        if stats is not None:
            stats["map"] = res["stats"]; stats["wall_ms"] = wall_
        o["summary"] = res["summary"]; o["gpu_ms"] += res["gpu_ms"]; o["launches"] += res["launches"]; o["d2h_bytes"] += m["d2h_bytes"]
        return o

    def run_pipelined(n_steps, e2e, stats=None):
        """n_steps steps over n_flight worker threads (worker w takes steps w, w + n_flight, ...); returns the last step's output."""
        turn[0] = 0
        outs = [None] * n_steps; errs = []

        def worker(w):
            try:
                mine = list(range(w, n_steps, n_flight))
                if e2e and mine:
                    wctx[w].stage_reads(0, r_host.data_ptr(), r_off)
                for j, i in enumerate(mine):
                    if e2e and j + 1 < len(mine):
                        wctx[w].stage_reads((j + 1) & 1, r_host.data_ptr(), r_off)
                    outs[i] = pipelined_step(i, w, (j & 1) if e2e else None, stats if i == n_steps - 1 else None)
            except Exception as e:      # release the ticket so that the other worker does not wait for ever
                errs.append(e)
                with cond:
                    turn[0] = 1 << 60; cond.notify_all()
        ths = [threading.Thread(target=worker, args=(w,)) for w in range(n_flight)]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        if errs:
            raise errs[0]
        return outs

    sampler = ClockSampler(local); sampler.start()
    stats = {}
    gpu_ms = 0.0; launches = 0
    if n_flight > 1:
        run_pipelined(n_flight, False)                 # warm the worker contexts (allocations)
    barrier()
    t0 = time.perf_counter()
    if n_flight > 1:
        outs = run_pipelined(args.steps, False, stats)
        out = outs[-1]
        for o in outs:
            gpu_ms += o["gpu_ms"]; launches += o["launches"]
    else:
        for i in range(args.steps):
            out = step_dev(stats)
            gpu_ms += out["gpu_ms"]; launches += out["launches"]
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    tt = torch.tensor([wall], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    wall = float(tt.item())
    bases = int(out["summary"]["total_bases_mapped_reads"])
    tb = torch.tensor([bases], device=dev, dtype=torch.float64)
    if world > 1 and not by_contigs:               # contig shards: every rank already counts all reads
        dist.all_reduce(tb)
    total_bases = float(tb.item())
    value = total_bases * args.steps / 1e6 / wall

    # e2e leg: host-buffer C-ABI calls.  Every step copies its reads from pinned host memory (mm_stage_reads_async, double-
    # buffered: the copy of a worker's next step runs while its current step computes) and reads every result array back; all
    # inside the timed region.
    def e2e_step(i, last):
        if by_contigs:
            return pipeline.map_and_classify_sharded(ctx, [ix], host_ptr=r_host.data_ptr(), offsets=r_off, read_len=read_len, contig_len=contig_len,
                                                     contig_taxon=contig_taxon, n_taxa=n_taxa, perc_identity=PI, min_read_len=wl["min_read_len"],
                                                     exchange=exchange, read_range=my_reads)
        if not last:
            ctx.stage_reads((i + 1) & 1, r_host.data_ptr(), r_off)
        return pipeline.map_and_classify(ctx, ix, staged_slot=i & 1, offsets=r_off, read_len=read_len, contig_len=contig_len,
                                         contig_taxon=contig_taxon, n_taxa=n_taxa, perc_identity=PI, min_read_len=wl["min_read_len"])

    def e2e_run(n):
        if n_flight > 1:
            return run_pipelined(n, True)[-1]
        if not by_contigs:
            ctx.stage_reads(0, r_host.data_ptr(), r_off)
        for i in range(n):
            o = e2e_step(i, i + 1 == n)
        return o
    e2e_run(2)
    barrier()
    t0 = time.perf_counter()
    o2 = e2e_run(args.steps)
    torch.cuda.synchronize()
    wall2 = time.perf_counter() - t0
    tt = torch.tensor([wall2], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    wall2 = float(tt.item())
    e2e_value = total_bases * args.steps / 1e6 / wall2
    sampler.stop_flag = True; sampler.join(timeout=2)

    if args.e2e_breakdown and not by_contigs and n_flight == 1 and rank == 0:
        sd = {}; step_dev(sd)
        print("[breakdown] resident:", json.dumps({"stage_ms": {k_: round(v, 2) for k_, v in sd["map"].items() if k_.endswith("_ms")},
                                                   "wall_ms": {k_: round(v, 2) for k_, v in sd["wall_ms"].items()}}), file=sys.stderr)
        ctx.stage_reads(0, r_host.data_ptr(), r_off)
        for i in range(3):
            se = {}
            ta = time.perf_counter()
            ctx.stage_reads((i + 1) & 1, r_host.data_ptr(), r_off)
            tb_ = time.perf_counter()
            pipeline.map_and_classify(ctx, ix, staged_slot=i & 1, offsets=r_off, read_len=read_len, contig_len=contig_len, contig_taxon=contig_taxon,
                                      n_taxa=n_taxa, perc_identity=PI, min_read_len=wl["min_read_len"], stats=se)
            tc = time.perf_counter()
            print("[breakdown] e2e step %d:" % i, json.dumps({"stage_call_ms": round((tb_ - ta) * 1e3, 2), "step_ms": round((tc - ta) * 1e3, 2),
                  "stage_ms": {k_: round(v, 2) for k_, v in se["map"].items() if k_.endswith("_ms")},
                  "wall_ms": {k_: round(v, 2) for k_, v in se["wall_ms"].items()}}), file=sys.stderr)
        torch.cuda.synchronize()

    if n_flight > 1:        # per-kernel event times of overlapped steps include the other batch's kernels: time one step alone
        stats = {}
        step_dev(stats)
    ms = stats["map"]
    # per-stage device time (CUDA events on the library's stream, last timed step) and the two heaviest single kernels,
    # each timed alone by its own event pair; algorithmic bytes per SURVEY.md 8(d) / DESIGN.md section 4
    stage_keys = ("sketch_ms", "read_sketch_ms", "l1_probe_ms", "l1_sort_ms", "l1_candidates_ms", "l2_setup_ms", "l2_classify_ms",
                  "l2_sweep_ms", "l2_strand_ms", "accept_ms")
    stage_ms = {k_: ms[k_] for k_ in stage_keys}
    kernels = {"sketch_blockmin_kernel (K1)": (ms["k1_kernel_ms"], ms["bases"] / 4 + 8 * ms["read_minimizers"]),
               "l2_sweep_band_kernel (K5b)": (ms["sweep_kernel_ms"], 2 * 8 * ms["span_elems"] + 32 * ms["sweep_items"])}
    dom = max(kernels, key=lambda k_: kernels[k_][0])
    dom_ms, dom_bytes = kernels[dom]
    peaks = {}; traffic = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    try:            # dram__bytes_read.sum + dram__bytes_write.sum per launch, from the committed ncu --set full capture
        traffic = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    line = {
        "metric": "long_read_Mbp_per_s_map_plus_EM_classify", "value": value, "unit": "Mbp/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": wall / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32/int64 (mapping), f64 (mapq, EM)", "data": "synthetic",
        "config": {"workload": args.workload, "reads_per_gpu": wl["n_reads"], "db_gbp": istats["n_contigs"] * wl["contig_len"] / 1e9,
                   "k": K, "w": wl["w"], "min_read_len": wl["min_read_len"], "perc_identity": PI, "parallelism": (f"index sharded by contig range x{world}, every rank maps all {world}x{wl['n_reads']} reads, mappings all-gathered"
                                   if by_contigs else f"reads sharded x{world}, index replicated"),
                   "l2_flush": "inputs (index %.1f GB + reads) larger than L2" % (istats["device_bytes"] / 1e9),
                   "batches_in_flight": n_flight,
                   "index_minimizers": istats["n_minimizers"], "index_build_s": index_s, "setup_s": setup_s,
                   "mappings_per_step": int(out["summary"]["n_mappings"]), "candidates_per_step": int(out["summary"]["n_candidates"]),
                   "em_iters": int(out["em"]["iters"]) if out["em"] else 0, "smem_swept": ms["smem_swept"], "ambiguous_reads": ms["ambiguous_reads"],
                   "span_elems": ms["span_elems"], "hits": ms["hits"], "sketch_elems": ms["sketch_elems"],
                   "wall_ms_last_step": {k_: round(v, 2) for k_, v in stats.get("wall_ms", {}).items()}},
        "clocks": sampler.result(),
        "e2e": {"value": e2e_value, "unit": "Mbp/s", "h2d_bytes_per_step": int(r_host.numel() + r_off.nbytes),
                "d2h_bytes_per_step": int(o2["d2h_bytes"])},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak if peak else None,
                     "traffic": (traffic.get(dom, {}).get("dram_bytes_per_launch") if args.workload == traffic.get("workload") else None),
                     "algorithmic_bytes_per_launch": dom_bytes, "kernel_ms_per_launch": dom_ms,
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md, of fallback)",
                     "note": "both kernels are latency/issue-bound integer work (DESIGN.md section 4): the fraction is reported as measured; "
                             "kernel and stage times are CUDA events of one step run alone right after the timed region (with two batches in "
                             "flight the events of a step also cover the other batch's kernels)",
                     "kernels": {k_: {"ms": v[0], "algorithmic_GBps": (v[1] / (v[0] * 1e-3) / 1e9 if v[0] > 0 else 0.0)} for k_, v in kernels.items()},
                     "stage_ms": stage_ms, "kernel_ms_per_step": gpu_ms / args.steps},
    }
    if rank == 0 and not args.no_cpu_baseline:
        try:
            from oracle import pyoracle
            if os.path.exists(pyoracle.REF_BIN):
                d = tempfile.mkdtemp(prefix="mmref_")
                fa, fq, rb, nc = reference_sample(wl, args.ref_contigs, args.ref_reads, d)
                threads = os.cpu_count() or 1
                t, t_index = run_reference_once(pyoracle.REF_BIN, d, wl, threads)
                line["cpu_baseline"] = {"value": rb / 1e6 / t, "unit": "Mbp/s", "cores": threads, "kind": "reference",
                                        "sample": f"{args.ref_reads} reads vs first {nc} contigs ({nc * wl['contig_len'] / 1e6:.0f} Mbp) of the DB recipe; "
                                                  f"reference CLI map (index build {t_index:.1f} s excluded) + classify, -t {threads}"}
        except Exception as e:          # the baseline is a reported number, never a reason to lose the bench line
            line["cpu_baseline"] = {"value": None, "unit": "Mbp/s", "cores": os.cpu_count(), "kind": "reference", "sample": "failed: %s" % e}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
