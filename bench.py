#!/usr/bin/env python
"""bench.py -- long-read Mbp/s, map + EM-classify (BASELINE.json metric), on N B200s of one node.

A "step" is one pass of the hot path (K0 pack -> K1 sketch -> K3 read sketch -> K4 L1 -> K5 L2 -> identity -> K6 mapq -> nLoc ->
K7/K8 EM; everything after the reads' upload stays in HBM) over one batch of synthetic reads against a GPU-resident index of a
synthetic DB.  At N = 1 the workload is BASELINE.json configs[1] ("config2": 100 k reads x 12 Gbp).

  value   reads already resident in HBM as ASCII when the timed region starts (mm_map_batch_dev + mm_classify_*)
  e2e     the same metric through the host-buffer C-ABI calls (pinned host reads staged with mm_stage_reads_async, every
          result array read back), H2D and D2H inside the timed region
N > 1 (torchrun): `value` / `e2e` = the index replicated, reads sharded (weak scaling: every rank maps its own batch), the only
collective is the per-round NCCL all-reduce of the EM taxon sums.  The same line also carries, under "extra":
  shard_contigs   north_star's split: the index sharded by contig range, every rank maps all N batches against its shard, the
                  accepted mappings are all-gathered and merged on the device (mm_classify_exchange), same total reads
  config3         BASELINE.json configs[2]: 1 M reads x 12 Gbp, strong scaling (the 1 M reads are split over the ranks), one EM
                  over all mappings; at N > 1 in both sharding modes
  config4_em      (N = 1) configs[3]: 200 near-identical strains, 50 EM rounds: EM ms/round and GB/s against 12 A + 4 R + 16 T
  same_config     (N = 1) the GPU CLI (metamaps_b200/metamaps) and the reference CLI on the SAME files (the cpu_baseline
                  sample): wall seconds of both, their ratio, and whether the ten output files are identical

--impl reference: the unmodified reference CLI built with the Boost shim (oracle/_ref/metamaps), all host threads, each step a
bounded sample of the same workload (cpu_baseline.sample in the output line says which).
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
    # BASELINE.json configs[3]: 500k reads vs 200 near-identical strains (0.1-0.5 % from one ancestor), 50 EM rounds
    "config4": dict(n_species=1, n_strains=200, contig_len=4_000_000, div=(0.001, 0.005), n_reads=500_000, mean_len=8000,
                    sigma=0.5, min_read_len=2000, w=16, seed=13, batch=5_000, em_rounds=50),
    # BASELINE.json configs[4] in miniature: the reference does not stay resident -- it is walked as contig-range chunks, one
    # resident per GPU at a time (the --maxmemory analogue), rank g of N owning chunks g, g+N, ...; thresholds follow the
    # reference's chunk chain (non-reset histogram); every chunk maps ALL reads; tables merged by (read, contig) on the device
    "config5-slice": dict(n_species=1000, n_strains=3, contig_len=4_000_000, div=0.01, n_reads=100_000, mean_len=8000,
                          sigma=0.5, min_read_len=2000, w=16, seed=11, chunk_contigs=125),
    "config5-tiny": dict(n_species=16, n_strains=3, contig_len=500_000, div=0.01, n_reads=2_000, mean_len=8000,
                         sigma=0.5, min_read_len=2000, w=16, seed=11, chunk_contigs=6),
    "config4-small": dict(n_species=1, n_strains=40, contig_len=1_000_000, div=(0.001, 0.005), n_reads=4_000, mean_len=8000,
                          sigma=0.5, min_read_len=2000, w=16, seed=13, batch=1_000, em_rounds=50),
}
K = 16
PI = 80.0
METRIC = "long_read_Mbp_per_s_map_plus_EM_classify"


def static_config(name, wl):
    """The workload description both arms print verbatim (no measured values in here)."""
    nsp, nst, L = wl["n_species"], wl["n_strains"], wl["contig_len"]
    div = wl["div"]
    return {"workload": name,
            "db": f"{nsp} species x {nst} strains x {L / 1e6:g} Mbp = {nsp * nst * L / 1e9:g} Gbp, strain divergence "
                  + (f"{div[0] * 100:g}-{div[1] * 100:g} %" if isinstance(div, tuple) else f"{div * 100:g} %") + f", seed {wl['seed']}",
            "reads": f"{wl['n_reads']} per GPU and step, log-normal mean {wl['mean_len']} b sigma {wl['sigma']}, 12 % sub/ins/del errors",
            "k": K, "w": wl["w"], "min_read_len": wl["min_read_len"], "perc_identity": PI,
            "commands": "mapDirectly --all + classify",
            "l2_flush": "inputs larger than L2 (index + reads of a step >> 126 MB); every step maps the same batch"}


# ------------------------------------------------------------------------------------------ synthetic data
def gen_db(torch, dev, wl):
    """ASCII DB on the device: n_species ancestors, each strain = ancestor with `div` substitutions (a (lo, hi) tuple draws the
    divergence of every strain uniformly: config 4's star of near-identical strains)."""
    g = torch.Generator(device=dev); g.manual_seed(wl["seed"])
    n_contigs = wl["n_species"] * wl["n_strains"]; L = wl["contig_len"]
    asc = torch.empty(n_contigs * L, dtype=torch.uint8, device=dev)
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
    codes = torch.empty(n_contigs * L, dtype=torch.uint8, device=dev)
    drng = np.random.Generator(np.random.PCG64(wl["seed"] + 5))
    ci = 0
    for s in range(wl["n_species"]):
        anc = torch.randint(0, 4, (L,), dtype=torch.uint8, device=dev, generator=g)
        for t in range(wl["n_strains"]):
            d = wl["div"] if not isinstance(wl["div"], tuple) else float(drng.uniform(*wl["div"]))
            mut = torch.rand(L, device=dev, generator=g) < d
            sub = torch.randint(1, 4, (L,), dtype=torch.uint8, device=dev, generator=g)
            c = torch.where(mut, (anc + sub) & 3, anc)
            codes[ci * L:(ci + 1) * L] = c
            asc[ci * L:(ci + 1) * L] = lut[c.long()]
            ci += 1
    offsets = np.arange(n_contigs + 1, dtype=np.int64) * L
    contig_taxon = np.arange(n_contigs, dtype=np.int32)          # one taxon (strain) per contig
    contig_len = np.full(n_contigs, L, np.int64)
    return asc, codes, offsets, contig_taxon, contig_len


def gen_reads(torch, dev, wl, codes, block, n=None, err=0.12):
    """Reads sampled from the DB with 12 % sub/ins/del errors (simulate.pl:57), random strand; ASCII on the device."""
    rng = np.random.Generator(np.random.PCG64(wl["seed"] * 1000 + block))
    n = n or wl["n_reads"]; L = wl["contig_len"]; n_contigs = wl["n_species"] * wl["n_strains"]
    mu = np.log(wl["mean_len"]) - 0.5 * wl["sigma"] ** 2
    lens = np.clip(rng.lognormal(mu, wl["sigma"], n), 1200, 40000).astype(np.int64)
    contig = rng.integers(0, n_contigs, n); start = (rng.random(n) * (L - lens)).astype(np.int64)
    rev = rng.random(n) < 0.5
    g = torch.Generator(device=dev); g.manual_seed(wl["seed"] * 7919 + block)
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
    db = synth.make_db(wl["seed"], sp, wl["n_strains"], wl["contig_len"], wl["div"] if not isinstance(wl["div"], tuple) else wl["div"][1])
    fa = synth.write_db(db, os.path.join(outdir, "db"))
    names, reads, _ = synth.make_reads(db, wl["seed"] + 1, n_reads, wl["mean_len"], lognormal_sigma=wl["sigma"])
    fq = os.path.join(outdir, "reads.fq")
    synth.write_fastq(fq, names, reads)
    bases = int(sum(len(r) for r in reads if len(r) >= wl["min_read_len"]))
    return fa, fq, bases, len(db.contig_codes)


def run_cli_once(binary, d, wl, threads, out="out", timing=None):
    """mapDirectly + classify with a `metamaps` CLI (the reference's, or this repo's).  Returns (seconds spent mapping +
    classifying, seconds of index build, wall seconds of both commands).  The reference builds its index inside mapDirectly;
    its own log line gives the mapping time."""
    import re
    os.makedirs(os.path.join(d, out), exist_ok=True)
    env = dict(os.environ, MM_HOST_TIMING="1") if timing is not None else None
    t0 = time.time()
    p = subprocess.run([binary, "mapDirectly", "--all", "-r", "db/DB.fa", "-q", "reads.fq", "-o", out + "/ref", "-m", str(wl["min_read_len"]),
                        "-w", str(wl["w"]), "-t", str(threads)], cwd=d, capture_output=True, text=True, env=env)
    t_map_total = time.time() - t0
    if p.returncode != 0:
        raise RuntimeError(os.path.basename(binary) + " mapDirectly failed: " + p.stderr[-500:])
    m = re.search(r"Time spent mapping the query : ([0-9.eE+-]+) sec", p.stdout)
    t_map = float(m.group(1)) if m else t_map_total
    t1 = time.time()
    if timing is not None:
        timing.append("mapDirectly %.3f s: " % t_map_total + "; ".join(l[len("[host timing] "):] for l in p.stderr.splitlines() if l.startswith("[host timing]")))
    p = subprocess.run([binary, "classify", "--DB", "db", "--mappings", out + "/ref", "-t", str(threads)], cwd=d, capture_output=True, text=True, env=env)
    t_cls = time.time() - t1
    if timing is not None:
        timing.append("classify %.3f s: " % t_cls + "; ".join(l[len("[host timing] "):] for l in p.stderr.splitlines() if l.startswith("[host timing]")))
    if p.returncode != 0:
        raise RuntimeError(os.path.basename(binary) + " classify failed: " + p.stderr[-500:])
    return t_map + t_cls, t_map_total - t_map, t_map_total + t_cls


def sample_text(args, wl, nc, threads):
    return (f"{args.ref_reads} reads (mean {wl['mean_len']} b, log-normal) vs the first {nc} contigs ({nc * wl['contig_len'] / 1e6:.0f} Mbp) of the "
            f"{args.workload} DB recipe; reference CLI mapDirectly (index build excluded, its own 'Time spent mapping' line) + classify, -t {threads}")


def reference_arm(args, wl):
    from oracle import pyoracle
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    base = {"impl": "reference", "metric": METRIC, "unit": "Mbp/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32/int64 (mapping), f64 (mapq, EM)", "data": "synthetic", "config": static_config(args.workload, wl)}
    if not os.path.exists(pyoracle.REF_BIN):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/metamaps missing (build it where /root/reference exists)"}))
        return
    d = tempfile.mkdtemp(prefix="mmref_")
    fa, fq, bases, nc = reference_sample(wl, args.ref_contigs, args.ref_reads, d)
    times = []
    for i in range(args.warmup + args.steps):
        t, t_index, _ = run_cli_once(pyoracle.REF_BIN, d, wl, threads)
        if i >= args.warmup:
            times.append(t)
    sec = float(np.mean(times))
    val = bases / 1e6 / sec
    sample = sample_text(args, wl, nc, threads)
    base.update({"value": val, "ms_per_step": sec * 1e3,
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
    ap.add_argument("--reads", type=int, default=0, help="override reads per GPU and step")
    ap.add_argument("--ref-contigs", type=int, default=12, help="reference arm / cpu_baseline: DB sample size in contigs")
    ap.add_argument("--ref-reads", type=int, default=1500, help="reference arm / cpu_baseline: reads in the sample")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the cpu_baseline and same_config legs")
    ap.add_argument("--no-extras", action="store_true", help="skip the extra legs (shard_contigs, config3, config4_em)")
    ap.add_argument("--config4-reads", type=int, default=100_000, help="reads of the config-4 EM leg inside the default run (the workload itself: --workload config4)")
    ap.add_argument("--shard", default="reads", choices=["reads", "contigs"],
                    help="N > 1, what `value` measures: 'reads' = index replicated, every rank maps its own reads; 'contigs' = the index is "
                         "split into contig ranges, every rank maps ALL reads against its shard, mappings are exchanged on the device")
    ap.add_argument("--e2e-breakdown", action="store_true", help="per-stage CUDA-event times and wall-clock split of e2e steps on stderr")
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
    # torchrun exports OMP_NUM_THREADS=1; the library's few host loops (table building) share the node's cores between the ranks
    os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 1) // world))
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

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t)
        return float(t.item())

    if args.workload.startswith("config4"):
        line = config4_run(args, wl, torch, dev, ctx, capi, pipeline, full=True)
        if rank == 0:
            print(json.dumps(line))
        return

    if args.workload.startswith("config5"):
        line = config5_run(args, wl, torch, dist, dev, ctx, capi, pipeline, world, rank, barrier, max_over_ranks)
        if rank == 0:
            print(json.dumps(line))
        if world > 1:
            dist.destroy_process_group()
        return

    t_setup = time.time()
    asc, codes, offsets, contig_taxon, contig_len = gen_db(torch, dev, wl)
    torch.cuda.synchronize()
    n_contigs = len(offsets) - 1
    n_taxa = int(contig_taxon.max()) + 1
    step_c = max(1, (256_000_000 // wl["contig_len"]))

    def build_index(c_lo, c_hi, shard):
        ix = capi.Index(ctx, K, wl["w"])
        if shard:
            ix.set_shard(c_lo, keep_counts=True)
        for c0 in range(c_lo, c_hi, step_c):
            c1 = min(c_hi, c0 + step_c)
            ix.add_dev(asc.data_ptr(), offsets[c0:c1 + 1])
        ix.finalize()
        if shard:
            ix.sync_threshold()                        # collective: occurrence threshold of the whole reference
        return ix

    by_contigs = args.shard == "contigs" and world > 1
    t_ix = time.time()
    own = (rank * n_contigs // world, (rank + 1) * n_contigs // world)
    ix = build_index(own[0], own[1], True) if by_contigs else build_index(0, n_contigs, False)
    index_s = time.time() - t_ix
    istats = ix.stats()

    def all_blocks_reads(n_blocks, first_block=0, n=None):
        """The reads of blocks first_block .. +n_blocks concatenated (contig-sharded ranks all map the same reads)."""
        blocks = [gen_reads(torch, dev, wl, codes, first_block + b, n) for b in range(n_blocks)]
        r_asc_ = torch.cat([b[0] for b in blocks]) if n_blocks > 1 else blocks[0][0]
        off_ = [np.zeros(1, np.int64)]; base = 0
        for b in blocks:
            off_.append(b[1][1:] + base); base += int(b[1][-1])
        cnt = np.cumsum([0] + [len(b[1]) - 1 for b in blocks])
        return r_asc_, np.concatenate(off_).astype(np.int64), cnt

    if by_contigs:                                 # every rank needs every read: blocks 0..N-1 of the read recipe
        r_asc, r_off, blk = all_blocks_reads(world)
        my_reads = (int(blk[rank]), int(blk[rank + 1]))
    else:
        r_asc, r_off = gen_reads(torch, dev, wl, codes, rank)
        my_reads = None
    torch.cuda.empty_cache()
    read_len = np.diff(r_off).astype(np.int32)
    r_host = torch.empty(r_asc.numel(), dtype=torch.uint8, pin_memory=True)      # pinned host copy of the reads for the e2e leg
    r_host.copy_(r_asc); torch.cuda.synchronize()
    ctx.classify_setup(contig_len, contig_taxon, n_taxa)
    setup_s = time.time() - t_setup
    common = dict(contig_len=contig_len, contig_taxon=contig_taxon, n_taxa=n_taxa, perc_identity=PI, min_read_len=wl["min_read_len"])

    def step_dev(stats=None):
        if by_contigs:
            return pipeline.map_and_classify_sharded(ctx, [ix], dev_ptr=r_asc.data_ptr(), offsets=r_off, read_len=read_len, read_range=my_reads, stats=stats, **common)
        return pipeline.map_and_classify(ctx, ix, dev_ptr=r_asc.data_ptr(), offsets=r_off, read_len=read_len, stats=stats, **common)

    for _ in range(args.warmup):
        out = step_dev()
    if args.profile_step:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        step_dev()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return

    sampler = ClockSampler(local); sampler.start()
    stats = {}
    gpu_ms = 0.0; launches = 0
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        out = step_dev(stats)
        gpu_ms += out["gpu_ms"]; launches += out["launches"]
    torch.cuda.synchronize()
    wall = max_over_ranks(time.perf_counter() - t0)
    bases = int(out["summary"]["total_bases_mapped_reads"])
    total_bases = float(bases) if by_contigs else sum_over_ranks(bases)       # contig shards: every rank already counts all reads
    value = total_bases * args.steps / 1e6 / wall

    # e2e leg: host-buffer C-ABI calls.  Every step copies its reads from pinned host memory (mm_stage_reads_async, double-
    # buffered: the upload of the next step runs while the current step computes) and reads every result array back; all
    # inside the timed region.
    def e2e_step(i, last, st=None):
        if by_contigs:
            return pipeline.map_and_classify_sharded(ctx, [ix], host_ptr=r_host.data_ptr(), offsets=r_off, read_len=read_len, read_range=my_reads, **common)
        if not last:
            ctx.stage_reads((i + 1) & 1, r_host.data_ptr(), r_off)
        return pipeline.map_and_classify(ctx, ix, staged_slot=i & 1, offsets=r_off, read_len=read_len, stats=st, **common)

    def e2e_run(n):
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
    wall2 = max_over_ranks(time.perf_counter() - t0)
    e2e_value = total_bases * args.steps / 1e6 / wall2
    sampler.stop_flag = True; sampler.join(timeout=2)

    if args.e2e_breakdown and not by_contigs and rank == 0:
        sd = {}; step_dev(sd)
        print("[breakdown] resident:", json.dumps({"stage_ms": {k_: round(v, 2) for k_, v in sd["map"].items() if k_.endswith("_ms")},
                                                   "wall_ms": {k_: round(v, 2) for k_, v in sd["wall_ms"].items()}}), file=sys.stderr)
        ctx.stage_reads(0, r_host.data_ptr(), r_off)
        for i in range(3):
            se = {}
            ta = time.perf_counter()
            e2e_step(i, False, se)
            tc = time.perf_counter()
            print("[breakdown] e2e step %d:" % i, json.dumps({"step_ms": round((tc - ta) * 1e3, 2),
                  "stage_ms": {k_: round(v, 2) for k_, v in se["map"].items() if k_.endswith("_ms")},
                  "wall_ms": {k_: round(v, 2) for k_, v in se["wall_ms"].items()}}), file=sys.stderr)
        torch.cuda.synchronize()

    # ---- roofline: per-kernel table (last timed step; CUDA events on the library's stream) against SURVEY.md 8(d)'s bytes
    ms = stats["map"]; cs = out["classify"]
    peaks = {}; traffic = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    for fn in ("r2_traffic.json", "r1_traffic.json"):      # dram__bytes_read.sum + dram__bytes_write.sum per launch, committed ncu --set full captures
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", fn)))
            break
        except Exception:
            pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    sm_clock_hz = float(peaks.get("sm_max_mhz", 1965.0)) * 1e6
    A = int(cs["n_mappings"]); R_ = int(cs["n_reads_mapped"]); rounds = int(cs["em_iters"])
    s_tot, H, C_, span = ms["sketch_elems"], ms["hits"], ms["candidates"], ms["span_elems"]
    k5_bytes = 8 * span + 4 * s_tot + 20 * C_
    table = {
        "K1 sketch_blockmin_kernel": {"ms": ms["k1_kernel_ms"], "bytes": ms["bases"] / 4 + 8 * ms["read_minimizers"],
                                      # 73 IMAD-pipe instructions per position (two 64-bit Murmur3 hashes), 64 lanes/clk/SM: the roof that binds K1
                                      "imad_roof_ms": 73 * ms["bases"] / (ctx_sm_count(torch, dev) * 64 * sm_clock_hz) * 1e3},
        "K1 stage (pack + sketch + compaction)": {"ms": ms["sketch_ms"], "bytes": 1.25 * ms["bases"] + 8 * ms["read_minimizers"]},
        "K3 read sketch": {"ms": ms["read_sketch_ms"], "bytes": 2 * 8 * ms["read_minimizers"]},
        "K4 L1 (probe + filter + sort + loci)": {"ms": ms["l1_probe_ms"] + ms["l1_sort_ms"] + ms["l1_candidates_ms"], "bytes": 16 * s_tot + 8 * H + 12 * C_},
        "K5 L2 (setup + classify + prune + sweep + strand)": {"ms": ms["l2_setup_ms"] + ms["l2_classify_ms"] + ms["l2_prune_ms"] + ms["l2_sweep_ms"] + ms["l2_strand_ms"], "bytes": k5_bytes},
        "K4 l1_probe_filter_kernel": {"ms": ms["l1_kernel_ms"], "bytes": 16 * s_tot + 8 * H + 12 * C_},
        "K5a l2_classify_smem_kernel": {"ms": ms["l2_classify_ms"], "bytes": 8 * span},
        "K5p l2_prune_warp_kernel": {"ms": ms["l2_prune_ms"], "bytes": 16 * span / 32 + 8 * C_},
        "K5b l2_sweep_band_kernel": {"ms": ms["sweep_kernel_ms"], "bytes": k5_bytes},
        "K5c l2_strand_warp_kernel": {"ms": ms["l2_strand_ms"], "bytes": 8 * span / 2.8},
        "K6-K8 classify stage (identity, mapq, nLoc, EM)": {"ms": cs["classify_ms"], "bytes": 24 * A + (12 * A + 4 * R_ + 16 * n_taxa) * rounds + 8 * A},
    }
    for v in table.values():
        v["GBps"] = v["bytes"] / (v["ms"] * 1e-3) / 1e9 if v["ms"] > 0 else 0.0
        v["frac"] = v["GBps"] / peak
        if "imad_roof_ms" in v:
            v["imad_frac"] = v["imad_roof_ms"] / v["ms"] if v["ms"] > 0 else 0.0
    dom = max(("K1 sketch_blockmin_kernel", "K4 l1_probe_filter_kernel", "K5a l2_classify_smem_kernel", "K5b l2_sweep_band_kernel"), key=lambda k_: table[k_]["ms"])
    dom_ms, dom_bytes = table[dom]["ms"], table[dom]["bytes"]
    achieved = table[dom]["GBps"]
    step_bytes = sum(table[k_]["bytes"] for k_ in ("K1 stage (pack + sketch + compaction)", "K3 read sketch", "K4 L1 (probe + filter + sort + loci)",
                                                    "K5 L2 (setup + classify + prune + sweep + strand)", "K6-K8 classify stage (identity, mapq, nLoc, EM)"))
    detail = {"reads_per_gpu": wl["n_reads"], "db_gbp": n_contigs * wl["contig_len"] / 1e9,
              "parallelism": (f"index sharded by contig range x{world}, every rank maps all {world}x{wl['n_reads']} reads, mappings all-gathered on the device"
                              if by_contigs else f"reads sharded x{world}, index replicated"),
              "index_device_gb": istats["device_bytes"] / 1e9, "index_minimizers": istats["n_minimizers"], "index_build_s": index_s, "setup_s": setup_s,
              "mappings_per_step": int(out["summary"]["n_mappings"]), "candidates_per_step": int(out["summary"]["n_candidates"]),
              "em_iters": rounds, "identity_fixups": int(cs["n_identity_fixups"]), "smem_swept": ms["smem_swept"], "ambiguous_reads": ms["ambiguous_reads"],
              "span_elems": span, "hits": H, "hits_kept": ms["hits_kept"], "sketch_elems": s_tot, "sweep_items": ms["sweep_items"],
              "window_starts": ms["window_starts"], "window_starts_swept": ms["window_starts_swept"],
              "stage_ms": {k_: ms[k_] for k_ in ms if k_.endswith("_ms")}, "classify_ms": cs["classify_ms"], "em_ms": cs["em_ms"],
              "kernel_ms_per_step": gpu_ms / args.steps, "step_algorithmic_GB": step_bytes / 1e9,
              "step_algorithmic_frac_of_hbm": step_bytes / (wall / args.steps) / 1e9 / peak,
              "wall_ms_last_step": {k_: round(v, 2) for k_, v in stats.get("wall_ms", {}).items()}}
    line = {
        "metric": METRIC, "value": value, "unit": "Mbp/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": wall / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32/int64 (mapping), f64 (mapq, EM)", "data": "synthetic",
        "config": static_config(args.workload, wl),
        "clocks": sampler.result(),
        "e2e": {"value": e2e_value, "unit": "Mbp/s", "h2d_bytes_per_step": int(r_host.numel() + r_off.nbytes),
                "d2h_bytes_per_step": int(o2["d2h_bytes"])},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak if peak else None,
                     "traffic": (traffic.get(dom, {}).get("dram_bytes_per_launch") if args.workload == traffic.get("workload") else None),
                     "algorithmic_bytes_per_launch": dom_bytes, "kernel_ms_per_launch": dom_ms,
                     "algorithmic_bytes_formula": "SURVEY.md 8(d): K4 = 16*s + 8*H + 12*C (the whole L1 row, charged to the fused probe + filter kernel); K5 = 8*sum N_c + 4*s + 20*C "
                                                  "(the whole L2 row, charged to the sweep kernel); K1 = bases/4 + 8*n_min",
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md, of fallback)",
                     "note": "integer / latency-bound work: the HBM fraction is reported as measured; K1 is bound by the IMAD pipe (imad_frac = its share of that roof); "
                             "kernel and stage times are CUDA events of the last timed step",
                     "kernels": table},
        "detail": detail,
    }
    extra = {}
    del r_asc, r_host, out, o2
    holder = [ix]; del ix
    torch.cuda.empty_cache()
    if not args.no_extras:
        extra_legs(extra, args, wl, torch, dist, dev, ctx, capi, pipeline, world, rank, codes, asc, offsets, holder, by_contigs, build_index, own,
                   all_blocks_reads, common, barrier, max_over_ranks, sum_over_ranks, n_contigs)
    holder.clear()
    if rank == 0 and world == 1 and not args.no_extras and args.workload == "config2":
        try:
            extra["config4_em"] = config4_run(args, dict(WORKLOADS["config4"], n_reads=args.config4_reads), torch, dev, ctx, capi, pipeline, full=False)
        except Exception as e:
            extra["config4_em"] = {"error": "%s: %s" % (type(e).__name__, e)}
    if rank == 0 and not args.no_cpu_baseline:
        try:
            from oracle import pyoracle
            if os.path.exists(pyoracle.REF_BIN):
                d = tempfile.mkdtemp(prefix="mmref_")
                fa, fq, rb, nc = reference_sample(wl, args.ref_contigs, args.ref_reads, d)
                threads = os.cpu_count() or 1
                t, t_index, t_wall = run_cli_once(pyoracle.REF_BIN, d, wl, threads)
                line["cpu_baseline"] = {"value": rb / 1e6 / t, "unit": "Mbp/s", "cores": threads, "kind": "reference",
                                        "sample": sample_text(args, wl, nc, threads) + f" (index build {t_index:.1f} s)"}
                if world == 1:
                    extra["same_config"] = same_config_leg(d, wl, rb, t, t_index, t_wall, threads)
        except Exception as e:          # the baseline is a reported number, never a reason to lose the bench line
            line["cpu_baseline"] = {"value": None, "unit": "Mbp/s", "cores": os.cpu_count(), "kind": "reference", "sample": "failed: %s" % e}
    line["extra"] = extra
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def ctx_sm_count(torch, dev):
    return torch.cuda.get_device_properties(dev).multi_processor_count


# ------------------------------------------------------------------------------------------ same-config leg
def same_config_leg(d, wl, bases, ref_map_cls_s, ref_index_s, ref_wall_s, threads):
    """This repo's CLI (C++ host over the CUDA library) on the exact files the reference CLI just processed: wall seconds of
    both, and a byte / 1e-6 comparison of the ten output files."""
    from metamaps_b200 import build
    from tests import cli_common
    if not os.path.exists(build.HOST_BIN):
        return {"error": "metamaps_b200/metamaps not built"}
    run_cli_once(build.HOST_BIN, d, wl, threads, out="out_b200")          # warm-up: CUDA context creation, page cache
    host_timing = []
    g_map_cls, _, g_wall = run_cli_once(build.HOST_BIN, d, wl, threads, out="out_b200", timing=host_timing)
    ok = True; identical = 0; err = None
    try:
        identical = cli_common.compare_dirs(os.path.join(d, "out"), os.path.join(d, "out_b200"))
    except AssertionError as e:
        ok = False; err = str(e)[:300]
    # CLI throughput on a read file large enough to amortise process start-up: 20 000 reads of the same recipe against the same DB sample
    cli = {}
    try:
        from metamaps_b200 import synth
        d2 = os.path.join(d, "cli_e2e"); os.makedirs(d2, exist_ok=True)
        os.symlink(os.path.join(d, "db"), os.path.join(d2, "db"))
        sp = max(1, len(open(os.path.join(d, "db", "taxonInfo.txt")).read().splitlines()) // wl["n_strains"])
        db = synth.make_db(wl["seed"], sp, wl["n_strains"], wl["contig_len"], wl["div"] if not isinstance(wl["div"], tuple) else wl["div"][1])
        names, reads, _ = synth.make_reads(db, wl["seed"] + 2, 20_000, wl["mean_len"], lognormal_sigma=wl["sigma"])
        synth.write_fastq(os.path.join(d2, "reads.fq"), names, reads)
        b2 = int(sum(len(r) for r in reads if len(r) >= wl["min_read_len"]))
        fq_bytes = os.path.getsize(os.path.join(d2, "reads.fq"))
        del db, reads
        ph = []
        _, _, w2 = run_cli_once(build.HOST_BIN, d2, wl, threads, out="out_b200", timing=ph)
        cli = {"reads": 20_000, "fastq_bytes": fq_bytes, "seconds": w2, "Mbp_per_s": b2 / 1e6 / w2, "phases": ph,
               "what": "metamaps_b200/metamaps mapDirectly --all + classify, FASTQ + FASTA in, ten files out, process start to exit (parser, GPU and writer threads)"}
    except Exception as e:
        cli = {"error": "%s: %s" % (type(e).__name__, e)}
    return {"files": "the cpu_baseline sample's db/ + reads.fq, both CLIs: mapDirectly --all -m %d -w %d + classify" % (wl["min_read_len"], wl["w"]),
            "cli_e2e": cli,
            "gpu_cli_s": g_wall, "ref_s": ref_wall_s, "ratio": ref_wall_s / g_wall if g_wall > 0 else None,
            "ref_map_plus_classify_s": ref_map_cls_s, "ref_index_build_s": ref_index_s, "ref_threads": threads,
            "gpu_cli_Mbp_per_s": bases / 1e6 / g_wall, "ref_Mbp_per_s_whole_run": bases / 1e6 / ref_wall_s,
            "gpu_cli_phases": host_timing, "outputs_match": ok, "files_identical": identical, "files_compared": 10, "mismatch": err,
            "note": "wall clock of the two commands of each CLI, process start to exit (the GPU side includes CUDA context creation, FASTA parsing and its index build)"}


# ------------------------------------------------------------------------------------------ extra legs
def extra_legs(extra, args, wl, torch, dist, dev, ctx, capi, pipeline, world, rank, codes, asc, offsets, holder, by_contigs, build_index, own, all_blocks_reads,
               common, barrier, max_over_ranks, sum_over_ranks, n_contigs):
    """Fills `extra` leg by leg; a failing leg leaves its error string and the others still run (every rank takes the same path:
    the legs are collective)."""
    # ---- config 3: 10 x the reads of a step in total (1 M for config 2), strong scaling over the ranks, one EM over all mappings
    n_blocks = 16
    per_block = max(1, wl["n_reads"] * 10 // n_blocks)

    def config3(mode, ix):
        mine = list(range(n_blocks)) if mode == "contigs" else list(range(rank, n_blocks, world))
        data = [gen_reads(torch, dev, wl, codes, 100 + b, per_block) for b in mine]
        torch.cuda.empty_cache()
        tot_reads = per_block * n_blocks
        lo, hi = rank * tot_reads // world, (rank + 1) * tot_reads // world

        def run():
            ctx.classify_begin()
            nb = 0
            for (a, off) in data:
                if mode == "contigs":                # every rank sketches its slice of the block; the sketches are all-gathered
                    nr = len(off) - 1; b0, b1 = rank * nr // world, (rank + 1) * nr // world
                    res = capi.map_reads_sharded(ctx, ix, a.data_ptr() + int(off[b0]), off[b0:b1 + 1] - off[b0], PI, wl["min_read_len"])
                else:
                    res = capi.map_reads(ctx, ix, None, PI, wl["min_read_len"], dev_ptr=a.data_ptr(), offsets=off, fetch=False)
                ctx.classify_add(getattr(ix, "first_contig", 0)); ctx.classify_next_batch()
                nb += int(res["summary"]["total_bases_mapped_reads"])
            if mode == "contigs":
                ctx.classify_exchange(lo, hi)
            cs = ctx.classify_run(0)
            r = ctx.classify_fetch(cs, what=("f", "best", "mapped_reads", "posterior"))
            return nb, cs, r
        run()                                       # warm-up (allocations)
        barrier()
        t0 = time.perf_counter()
        nb, cs, r = run()
        torch.cuda.synchronize()
        sec = max_over_ranks(time.perf_counter() - t0)
        bases = float(nb) if mode == "contigs" else sum_over_ranks(nb)
        del data
        torch.cuda.empty_cache()
        return {"reads_total": tot_reads, "batches": n_blocks, "mode": ("index sharded by contig range, mappings exchanged on the device" if mode == "contigs"
                                                                       else "reads sharded, index replicated"),
                "seconds": sec, "value": bases / 1e6 / sec, "unit": "Mbp/s", "scaling": "strong", "em_iters": int(cs["em_iters"]),
                "mappings_this_rank": int(cs["n_mappings"]), "f_sum": float(np.sum(r["f"]))}

    def leg(name, fn):
        try:
            extra[name] = fn()
        except Exception as e:
            extra[name] = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}
        torch.cuda.empty_cache()
    def h2d_probe():
        """What the host side can deliver when all ranks upload at once (what bounds `e2e` at N = 8): every rank copies 1 GiB of
        pinned host memory to its GPU five times, all ranks together; GB/s per rank = bytes / the slowest rank's time."""
        nb = 1 << 30
        hbuf = torch.empty(nb, dtype=torch.uint8, pin_memory=True); dbuf = torch.empty(nb, dtype=torch.uint8, device=dev)
        dbuf.copy_(hbuf, non_blocking=True); torch.cuda.synchronize()
        barrier()
        t0 = time.perf_counter()
        for _ in range(5):
            dbuf.copy_(hbuf, non_blocking=True)
        torch.cuda.synchronize()
        sec = max_over_ranks(time.perf_counter() - t0)
        del hbuf, dbuf
        return {"GBps_per_rank_all_ranks_uploading": 5 * nb / sec / 1e9, "GBps_aggregate": 5 * nb * world / sec / 1e9, "ranks": world,
                "what": "cudaMemcpyAsync pinned host -> device, 1 GiB x 5 per rank, all ranks concurrently (DMA; the e2e leg's K0 reads the same pinned pages over PCIe)"}
    leg("h2d_probe", h2d_probe)
    if not by_contigs:
        leg("config3", lambda: config3("reads", holder[0]))
    if world == 1:
        return
    # ---- north_star's split: contig-range shards.  Same total reads per step as the read-sharded `value` (N batches).
    if not by_contigs:
        holder.clear()                              # the replicated index goes before the shard is built
        torch.cuda.empty_cache()
        try:
            holder.append(build_index(own[0], own[1], True))
        except Exception as e:
            extra["shard_contigs"] = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}
            return
    ix_sh = holder[0]

    def shard_leg():
        # every rank holds ITS block of the step's reads (block `rank` of the recipe, the batch of the read-sharded `value`); it
        # sketches that block once, the sketches are all-gathered, all reads are mapped against the rank's shard
        r_asc, r_off = gen_reads(torch, dev, wl, codes, rank)
        torch.cuda.empty_cache()
        r_host = torch.empty(r_asc.numel(), dtype=torch.uint8, pin_memory=True); r_host.copy_(r_asc); torch.cuda.synchronize()
        stage = torch.empty_like(r_asc)

        def step(host):
            if host:                                   # e2e: the block comes from pinned host memory every step
                stage.copy_(r_host, non_blocking=True); torch.cuda.current_stream().synchronize()
            return pipeline.map_and_classify_sharded(ctx, [ix_sh], my_block=((stage if host else r_asc).data_ptr(), r_off), **common)
        res = {}
        for name, host in (("value", False), ("e2e", True)):
            for _ in range(3):
                o = step(host)
            barrier()
            t0 = time.perf_counter()
            n_it = max(2, min(args.steps, 5)); per = []
            for _ in range(n_it):
                ts = time.perf_counter(); o = step(host); per.append(round((time.perf_counter() - ts) * 1e3, 1))
            torch.cuda.synchronize()
            sec = max_over_ranks(time.perf_counter() - t0)
            res[name + "_step_ms_rank0"] = per
            res[name] = float(o["summary"]["total_bases_mapped_reads"]) * n_it / 1e6 / sec
            res["ms_per_step" if not host else "e2e_ms_per_step"] = sec / n_it * 1e3
        res.update({"unit": "Mbp/s", "reads_per_step": int(o["summary"]["n_reads"]), "mappings_all_shards_this_rank": int(o["summary"]["n_mappings_this_rank_all_shards"]),
                    "wall_ms_last_step": {k_: round(v, 2) for k_, v in o.get("wall_ms", {}).items()},
                    "parallelism": f"index sharded by contig range x{world}; every rank sketches its own block of {len(r_off) - 1} reads once, the sketches are "
                                   "all-gathered (ncclAllGather), all reads are mapped against the rank's shard; accepted mappings all-gathered and merged on "
                                   "the device; each rank finalises its block; EM all-reduce per round"})
        return res
    leg("shard_contigs", shard_leg)
    leg("config3_shard_contigs", lambda: config3("contigs", ix_sh))


# ------------------------------------------------------------------------------------------ config 5 (streamed reference chunks)
def config5_run(args, wl, torch, dist, dev, ctx, capi, pipeline, world, rank, barrier, max_over_ranks):
    t0 = time.time()
    asc, codes, offsets, contig_taxon, contig_len = gen_db(torch, dev, wl)
    n_contigs = len(offsets) - 1; n_taxa = n_contigs
    cc = wl["chunk_contigs"]; n_chunks = (n_contigs + cc - 1) // cc
    r_asc, r_off = gen_reads(torch, dev, wl, codes, 0)                      # every rank maps the same reads
    del codes
    n = len(r_off) - 1
    lo, hi = rank * n // world, (rank + 1) * n // world

    def build_chunk(c):
        c0, c1 = c * cc, min(n_contigs, (c + 1) * cc)
        ix = capi.Index(ctx, K, wl["w"]); ix.set_shard(c0, keep_counts=False)
        ix.add_dev(asc.data_ptr(), offsets[c0:c1 + 1]); ix.finalize()
        return ix

    def all_gather(obj):
        if world == 1:
            return [obj]
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out
    common = dict(dev_ptr=r_asc.data_ptr(), offsets=r_off, contig_len=contig_len, contig_taxon=contig_taxon, n_taxa=n_taxa, perc_identity=PI,
                  min_read_len=wl["min_read_len"])

    st5 = {}

    def run():
        st5.clear()
        return pipeline.map_and_classify_streamed(ctx, build_chunk, list(range(rank, n_chunks, world)), n_chunks, all_gather, read_range=(lo, hi), stats=st5, **common)
    setup_s = time.time() - t0
    for _ in range(max(1, min(args.warmup, 2))):
        out = run()
    barrier()
    t1 = time.perf_counter()
    steps = max(1, min(args.steps, 5))
    for _ in range(steps):
        out = run()
    torch.cuda.synchronize()
    sec = max_over_ranks(time.perf_counter() - t1) / steps
    bases = float(out["summary"]["total_bases_mapped_reads"])
    thr = {}
    for part in all_gather(out["thresholds"]):
        thr.update(part)
    check = None
    if world > 1:                       # the same chunks walked by ONE process (the reference's --maxmemory loop): rank 0 compares its block of reads
        if rank == 0:
            ctx1 = capi.Context(dev.index)
            keep = {k_: np.array(out[k_]) for k_ in ("read", "seq", "pos", "shared", "sketch", "strand", "identity", "mapq")}

            def build1(c):
                c0, c1 = c * cc, min(n_contigs, (c + 1) * cc)
                ix = capi.Index(ctx1, K, wl["w"]); ix.set_shard(c0, keep_counts=False)
                ix.add_dev(asc.data_ptr(), offsets[c0:c1 + 1]); ix.finalize()
                return ix
            ref = pipeline.map_and_classify_streamed(ctx1, build1, list(range(n_chunks)), n_chunks, lambda o: [o], em_max_iter=-1, **common)
            sel = (ref["read"] >= lo) & (ref["read"] < hi)
            same = all(np.array_equal(keep[k_], ref[k_][sel]) for k_ in keep)
            check = {"one_process_chunk_walk_equals_rank0_block": bool(same), "mappings_compared": int(sel.sum()),
                     "thresholds_equal": sorted(ref["thresholds"].items()) == sorted(thr.items())}
        barrier()
    finite = sum(1 for v in thr.values() if v != 0x7fffffff)
    return {"metric": METRIC, "value": bases / 1e6 / sec, "unit": "Mbp/s", "n_gpus": world, "steps": steps, "warmup": max(1, min(args.warmup, 2)),
            "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32/int64 (mapping), f64 (mapq, EM)",
            "data": "synthetic", "config": static_config(args.workload, wl),
            "detail": {"chunks": n_chunks, "contigs_per_chunk": cc, "chunk_mbp": cc * wl["contig_len"] / 1e6, "chunks_per_gpu": (n_chunks + world - 1) // world,
                       "resident_chunks_per_gpu": 1, "ownership": "interleaved: rank g owns chunks g, g+N, ...",
                       "thresholds": "reference --maxmemory chain (histogram and threshold carried from chunk to chunk, winSketch.hpp:302-304,452-495)",
                       "chunks_with_finite_threshold": finite, "threshold_values": sorted(set(int(v) for v in thr.values())),
                       "reads": n, "mappings_this_rank": int(out["classify"]["n_mappings"]), "em_iters": int(out["classify"]["em_iters"]),
                       "setup_s": setup_s, "check": check, "map_stage_ms_last_step": {k_: round(v, 1) for k_, v in st5.get("map_sum_ms", {}).items()},
                       "timed_region": "per step: every chunk of this rank built from the device-resident DB text, settled, mapped against all reads, freed; "
                                       "mappings exchanged and merged by (read, contig) on the device; mapq, EM (all-reduce per round); one D2H"}}


# ------------------------------------------------------------------------------------------ config 4 (EM stress)
def config4_run(args, wl, torch, dev, ctx, capi, pipeline, full):
    """200 near-identical strains: every read maps to (almost) every strain, so the EM has ~200 mappings per read to weigh.
    Reads are mapped in batches into ONE classify table, then the EM runs em_rounds rounds over all of it."""
    t0 = time.time()
    asc, codes, offsets, contig_taxon, contig_len = gen_db(torch, dev, wl)
    n_contigs = len(offsets) - 1; n_taxa = n_contigs
    ix = capi.Index(ctx, K, wl["w"])
    step_c = max(1, (256_000_000 // wl["contig_len"]))
    for c0 in range(0, n_contigs, step_c):
        ix.add_dev(asc.data_ptr(), offsets[c0:min(n_contigs, c0 + step_c) + 1])
    ix.finalize()
    del asc
    ctx.classify_setup(contig_len, contig_taxon, n_taxa)
    n_batches = max(1, (wl["n_reads"] + wl["batch"] - 1) // wl["batch"])
    setup_s = time.time() - t0
    ctx.classify_begin()
    # untimed warm-up batch: scratch buffers grow to this workload's sizes (cudaMalloc / cudaFree inside a step cost more than the kernels)
    a, off = gen_reads(torch, dev, wl, codes, 199, min(wl["batch"], wl["n_reads"]))
    capi.map_reads(ctx, ix, None, PI, wl["min_read_len"], dev_ptr=a.data_ptr(), offsets=off, fetch=False)
    del a
    bases = 0; map_ms = 0.0; cand = 0; stage = {}
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    for b in range(n_batches):
        a, off = gen_reads(torch, dev, wl, codes, 200 + b, min(wl["batch"], wl["n_reads"] - b * wl["batch"]))
        res = capi.map_reads(ctx, ix, None, PI, wl["min_read_len"], dev_ptr=a.data_ptr(), offsets=off, fetch=False)
        map_ms += res["gpu_ms"]; bases += int(res["summary"]["total_bases_mapped_reads"]); cand += int(res["summary"]["n_candidates"])
        for k_, v in ctx.last_map_stats().items():
            if k_.endswith("_ms") or k_.startswith("window_starts"):
                stage[k_] = stage.get(k_, 0) + float(v)
        ctx.classify_add(0); ctx.classify_next_batch()
        del a
    torch.cuda.synchronize()
    map_s = time.perf_counter() - t1
    ctx.classify_run(wl["em_rounds"])                 # warm-up (allocations, attribute opt-ins)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    cs = ctx.classify_run(wl["em_rounds"])
    torch.cuda.synchronize()
    cls_s = time.perf_counter() - t2
    cs_free = ctx.classify_run(0)                     # the reference's own stopping rule (fEM.h:636): how many rounds it takes here
    r = ctx.classify_fetch(cs, what=("f",))
    A = int(cs["n_mappings"]); R_ = int(cs["n_reads_mapped"]); rounds = int(cs["em_iters"])
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0))
    except Exception:
        peak = 6650.0
    em_bytes_round = 12 * A + 4 * R_ + 16 * n_taxa
    ms_round = cs["em_ms"] / max(rounds, 1)
    out = {"workload": "config4" if full else "config4 (bounded: %d of 500000 reads)" % wl["n_reads"], "strains": n_taxa, "db_mbp": n_contigs * wl["contig_len"] / 1e6,
           "reads": wl["n_reads"], "read_batches": n_batches, "mappings": A, "mappings_per_read": A / max(R_, 1), "candidates": cand,
           "em_rounds": rounds, "em_ms_total": cs["em_ms"], "em_ms_per_round": ms_round,
           "em_algorithmic_bytes_per_round": em_bytes_round, "em_GBps": em_bytes_round / (ms_round * 1e-3) / 1e9 if ms_round > 0 else 0.0,
           "em_frac_of_hbm_peak": (em_bytes_round / (ms_round * 1e-3) / 1e9 / peak) if ms_round > 0 else 0.0,
           "rounds_to_reference_stopping_rule": int(cs_free["em_iters"]),
           "map_s": map_s, "map_kernel_ms": map_ms, "classify_s": cls_s, "value": bases / 1e6 / (map_s + cls_s), "unit": "Mbp/s",
           "setup_s": setup_s, "f_sum": float(np.sum(r["f"])), "map_stage_ms": {k_: round(v, 1) for k_, v in stage.items()}}
    if not full:
        return out
    return {"metric": METRIC, "value": out["value"], "unit": "Mbp/s", "n_gpus": 1, "steps": 1, "warmup": 0, "ms_per_step": (map_s + cls_s) * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32/int64 (mapping), f64 (mapq, EM)", "data": "synthetic",
            "config": static_config(args.workload, wl), "detail": out}


if __name__ == "__main__":
    main()
