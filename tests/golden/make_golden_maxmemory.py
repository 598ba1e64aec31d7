#!/usr/bin/env python
"""Golden fixture `ref_maxmemory/`: the UNMODIFIED reference CLI (oracle/_ref/metamaps) with `--maxmemory 1` on a 60 Mbp reference
that its own memory estimate (winSketch.hpp:165-178,284-329) cuts into three chunks.

  pins   the reference's chunk loop: chunk N's sequence ids restart at 0, a read's lines are concatenated in chunk order by
         unifyFiles (mapWrap.h:128-132), and -- the quirk -- the occurrence histogram and threshold are NOT reset between chunks
         (winSketch.hpp:302-304,452-495): planted repeats give every chunk the same own distribution, yet the reference reports
         thresholds 4, 7 and 23 for chunks 0, 1, 2 (map.log keeps its INFO lines and its "Call storeCurrentState with N" chunk sizes).
Inputs are regenerated from the seed by `maxmemory_sample` (numpy, deterministic); only the reference's outputs are committed.
Run in the build container (needs /root/reference); the reference takes ~4 minutes here.
"""
import gzip
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from metamaps_b200 import synth  # noqa: E402

MAP_ARGS = ["--all", "-w", "2"]


def maxmemory_sample(outdir, n_contigs=30, L=2_000_000, seed=77):
    """30 random contigs of 2 Mbp; five 25-base units planted ~270/210/165/123/99 times over all contigs and, inside every contig,
    ten more units three times each (so each chunk alone sees the same occurrence histogram); 300 reads of 3 kb."""
    rng = np.random.default_rng(seed)
    contigs = [rng.integers(0, 4, L, dtype=np.uint8) for _ in range(n_contigs)]
    heavy = [rng.integers(0, 4, 25, dtype=np.uint8) for _ in range(5)]
    for u, copies in zip(heavy, (90, 70, 55, 41, 33)):
        for _ in range(copies * 3):
            ci = int(rng.integers(0, n_contigs)); pos = int(rng.integers(0, L - 50)); contigs[ci][pos:pos + 25] = u
    for ci in range(n_contigs):
        for _ in range(10):
            u = rng.integers(0, 4, 25, dtype=np.uint8)
            for _ in range(3):
                pos = int(rng.integers(0, L - 50)); contigs[ci][pos:pos + 25] = u
    names = [f"C{i}|kraken:taxid|{i + 1}|x" for i in range(n_contigs)]
    tax = {"1": ("1", "no rank", "root")}
    for i in range(n_contigs):
        tax[str(i + 1)] = ("1", "species", f"sp{i}")
    db = synth.SynthDB(names, [str(i + 1) for i in range(n_contigs)], contigs, tax)
    synth.write_db(db, os.path.join(outdir, "db"))
    nm, reads, _ = synth.make_reads(db, seed + 1, 300, 3000, err=0.08)
    synth.write_fastq(os.path.join(outdir, "reads.fq"), nm, reads)
    return db


def main():
    from oracle import pyoracle
    pyoracle.build()
    tmp = tempfile.mkdtemp(prefix="mm_maxmem_")
    maxmemory_sample(tmp)
    out = os.path.join(tmp, "out"); os.makedirs(out)
    p = subprocess.run([pyoracle.REF_BIN, "mapDirectly"] + MAP_ARGS + ["--mm", "1", "-r", "db/DB.fa", "-q", "reads.fq", "-o", "out/ref", "-t", str(os.cpu_count() or 1)],
                       cwd=tmp, check=True, capture_output=True, text=True)
    keep = [l for l in (p.stdout + p.stderr).splitlines() if "storeCurrentState" in l or "computeFreqHist, With threshold" in l]
    assert sum("storeCurrentState" in l for l in keep) >= 3, keep
    dst = os.path.join(HERE, "ref_maxmemory")
    shutil.rmtree(dst, ignore_errors=True); os.makedirs(dst)
    with open(os.path.join(dst, "map.log"), "w") as f:
        f.write("\n".join(keep) + "\n")
    for fn in sorted(os.listdir(out)):
        if fn.startswith("ref"):              # the mapping stage's files: ref, ref.meta, ref.meta.unmappedReadsLengths, ref.parameters
            with open(os.path.join(out, fn), "rb") as f, gzip.GzipFile(os.path.join(dst, fn + ".gz"), "wb", mtime=0) as g:
                g.write(f.read())
    print("\n".join(keep))
    shutil.rmtree(tmp)


if __name__ == "__main__":
    main()
