#!/usr/bin/env python
"""Golden fixture `ref_config2/`: the UNMODIFIED reference CLI (oracle/_ref/metamaps, `make -C oracle`) on a
config-2-shaped sample that is large enough for the reference's occurrence threshold to be FINITE.

  workload   bench.py's config-2 recipe cut to its first 34 species x 3 strains x 4 Mbp = 408 Mbp (1 % divergence),
             1500 reads (log-normal, mean 8 kb, 12 % errors), `mapDirectly --all -m 2000 -w 16` + `classify`
  pins       computeFreqHist's finite threshold (winSketch.hpp:452-495; the reference's own INFO line is kept in
             map.log), the over-frequent-hash skip of doL1Mapping (computeMap.hpp:314), the saturated-hash-space L1
             filter, and all ten output files on config-2-shaped data.

Run in the build container (needs /root/reference).  The inputs are regenerated from the seeds by
`config2_sample` (numpy, deterministic), so only the reference's OUTPUTS are committed (gzip).
"""
import gzip
import os
import shutil
import subprocess
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from metamaps_b200 import synth  # noqa: E402

SEED = 11
N_SPECIES, N_STRAINS, CONTIG_LEN, DIV = 34, 3, 4_000_000, 0.01
N_READS, MEAN_LEN, SIGMA, MIN_READ_LEN, W = 1500, 8000, 0.5, 2000, 16
MAP_ARGS = ["--all", "-m", str(MIN_READ_LEN), "-w", str(W)]


def config2_sample(outdir, n_species=N_SPECIES, n_reads=N_READS):
    """db/ + reads.fq of the sample under `outdir`; returns (db, fasta, fastq)."""
    db = synth.make_db(SEED, n_species, N_STRAINS, CONTIG_LEN, DIV)
    fa = synth.write_db(db, os.path.join(outdir, "db"))
    names, reads, _ = synth.make_reads(db, SEED + 1, n_reads, MEAN_LEN, lognormal_sigma=SIGMA)
    fq = os.path.join(outdir, "reads.fq")
    synth.write_fastq(fq, names, reads)
    return db, fa, fq


def main():
    from oracle import pyoracle
    pyoracle.build()
    tmp = tempfile.mkdtemp(prefix="mm_cfg2_")
    t0 = time.time()
    config2_sample(tmp)
    print("inputs written in %.0f s" % (time.time() - t0), flush=True)
    out = os.path.join(tmp, "out"); os.makedirs(out)
    t0 = time.time()
    with open(os.path.join(out, "map.log"), "w") as lg:
        subprocess.run([pyoracle.REF_BIN, "mapDirectly"] + MAP_ARGS + ["-r", "db/DB.fa", "-q", "reads.fq", "-o", "out/ref", "-t", str(os.cpu_count() or 1)],
                       cwd=tmp, check=True, stdout=lg, stderr=subprocess.STDOUT)
    print("reference mapDirectly: %.0f s" % (time.time() - t0), flush=True)
    t0 = time.time()
    with open(os.path.join(out, "classify.log"), "w") as lg:
        subprocess.run([pyoracle.REF_BIN, "classify", "--DB", "db", "--mappings", "out/ref", "-t", str(os.cpu_count() or 1)],
                       cwd=tmp, check=True, stdout=lg, stderr=subprocess.DEVNULL)
    print("reference classify: %.0f s" % (time.time() - t0), flush=True)
    dst = os.path.join(HERE, "ref_config2")
    shutil.rmtree(dst, ignore_errors=True); os.makedirs(dst)
    for fn in sorted(os.listdir(out)):
        if fn.startswith("ref") or fn.endswith(".log"):
            with open(os.path.join(out, fn), "rb") as f, gzip.GzipFile(os.path.join(dst, fn + ".gz"), "wb", mtime=0) as g:
                g.write(f.read())
    print(open(os.path.join(out, "map.log")).read())
    shutil.rmtree(tmp)


if __name__ == "__main__":
    main()
