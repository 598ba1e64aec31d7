#!/usr/bin/env python
"""Generates the committed golden fixtures in tests/golden/.  Run in the build container (needs
/root/reference and the shim-built reference in oracle/_ref; `make -C oracle`).

  example_kat.tsv        from the reference's own MetaMaps_example_output.zip (.EM file, 2985 mappings):
                         read length, col 10 (identity), col 11 (shared), col 12 (sketch size), col 13
                         (corrected identity).  Pins the float/double arithmetic of map_stats.hpp:44-54,
                         computeMap.hpp:406-411 and mapWrap.h:313-319.
  boost_binomial.json    binomial pmf / upper quantile / sf from SciPy's Boost.Math-backed ufuncs
                         (scipy.special._ufuncs._binom_pmf/_binom_isf/_binom_sf): the pin for the Boost calls
                         the reference makes (map_stats.hpp:88,204; mapWrap.h:340).
  ref_small/             outputs of the UNMODIFIED reference CLI (oracle/_ref/metamaps) on the seeded
                         synthetic workload `small` (metamaps_b200.synth): mapping file, .meta, .EM*, plus
                         ref_small_internals.npz = per-read L1/L2 internals and minimizer streams obtained
                         through oracle/_ref/libmm_refharness.so (the reference's own headers).
"""
import json
import os
import shutil
import subprocess
import sys
import tempfile
import zipfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from metamaps_b200 import synth  # noqa: E402
from oracle import pyoracle  # noqa: E402


def small_workload(outdir):
    """The `small` golden workload: 4 species x 2 strains x 60 kbp, 200 reads of 3 kb, defaults (-m 1000)."""
    db = synth.make_db(21, 4, 2, 60_000, 0.02)
    fa = synth.write_db(db, os.path.join(outdir, "db"))
    names, reads, truth = synth.make_reads(db, 22, 200, 3000, frac_short=0.03, frac_random=0.02)
    fq = os.path.join(outdir, "reads.fq")
    synth.write_fastq(fq, names, reads)
    return db, fa, fq, names, reads


def main():
    # 1. example KAT
    z = zipfile.ZipFile("/root/reference/MetaMaps_example_output.zip")
    em = z.read("MetaMaps_example_output/hmp7_2_short_miniSeq+H.EM").decode().splitlines()
    with open(os.path.join(HERE, "example_kat.tsv"), "w") as f:
        f.write("#read_len\tcol10_identity\tcol11_shared\tcol12_sketch\tcol13_corrected_identity\n")
        for line in em:
            c = line.split(" ")
            f.write("\t".join([c[1], c[9], c[10], c[11], c[12]]) + "\n")
    params = z.read("MetaMaps_example_output/hmp7_2_short_miniSeq+H.parameters").decode()
    meta = z.read("MetaMaps_example_output/hmp7_2_short_miniSeq+H.meta").decode()
    with open(os.path.join(HERE, "example_parameters.txt"), "w") as f:
        f.write(params)
    with open(os.path.join(HERE, "example_meta.txt"), "w") as f:
        f.write(meta)

    # 2. Boost binomial vectors
    import scipy.special._ufuncs as u
    rng = np.random.default_rng(99)
    cases = []
    for _ in range(600):
        n = int(rng.integers(1, 6000))
        p = float(rng.choice([rng.uniform(1e-4, 0.05), rng.uniform(0.05, 0.6), rng.uniform(0.6, 0.999)]))
        k = int(np.clip(rng.normal(n * p, 3 * np.sqrt(n * p * (1 - p)) + 1), 0, n))
        q = float(rng.choice([0.05000000074505806, 0.05, 0.5, 1e-3, 0.9]))   # 0.0500000007 = float((1-0.9f)/2)
        cases.append({"n": n, "p": p, "k": k, "q": q,
                      "pmf": float(u._binom_pmf(float(k), float(n), p)),
                      "isf": float(u._binom_isf(q, float(n), p)),
                      "sf": float(u._binom_sf(float(k), float(n), p))})
    with open(os.path.join(HERE, "boost_binomial.json"), "w") as f:
        json.dump({"source": "scipy %s scipy.special._ufuncs (Boost.Math)" % __import__("scipy").__version__, "cases": cases}, f)

    # 3. reference outputs on the small workload
    pyoracle.build()
    tmp = tempfile.mkdtemp()
    db, fa, fq, names, reads = small_workload(tmp)
    out = os.path.join(tmp, "out")
    os.makedirs(out)
    subprocess.run([pyoracle.REF_BIN, "mapDirectly", "--all", "-r", "db/DB.fa", "-q", "reads.fq", "-o", "out/ref"],
                   cwd=tmp, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    subprocess.run([pyoracle.REF_BIN, "classify", "--DB", "db", "--mappings", "out/ref"],
                   cwd=tmp, check=True, stdout=open(os.path.join(out, "classify.log"), "w"), stderr=subprocess.DEVNULL)
    dst = os.path.join(HERE, "ref_small")
    os.makedirs(dst, exist_ok=True)
    for fn in os.listdir(out):
        if fn.startswith("ref") or fn == "classify.log":
            shutil.copy(os.path.join(out, fn), os.path.join(dst, fn))
    # internals through the reference's own headers
    R = pyoracle.RefHarness()
    k = 16
    w = int([l for l in open(os.path.join(out, "ref.parameters")) if l.startswith("windowSize")][0].split()[1])
    h = R.index_build_fasta(fa, k, w)
    ih, isq, iwp, ist = R.index_get(h)
    rec = {"w": w, "k": k, "index_hash": ih, "index_seq": isq, "index_wpos": iwp, "index_strand": ist,
           "index_unique": R.index_unique(h), "freq_threshold": R.index_freq_threshold(h)}
    s_list, mh_list, off = [], [], [0]
    keys = ["seq", "start", "end", "pos", "shared", "votes", "valid", "optStart", "optEnd"]
    acc = {kk: [] for kk in keys}
    q_off, q_hash, q_strand = [0], [], []
    for rd in reads:
        s = synth.codes_to_ascii(rd)
        if len(s) < 1000:
            s_list.append(0); mh_list.append(0); off.append(off[-1]); q_off.append(q_off[-1]); continue
        m = R.map_read(h, s)
        s_list.append(m["s"]); mh_list.append(m["minimumHits"]); off.append(off[-1] + len(m["seq"]))
        for kk in keys:
            acc[kk].append(m[kk])
        qh, qw, qs = R.read_sketch(h, s)
        q_hash.append(qh); q_strand.append(qs); q_off.append(q_off[-1] + len(qh))
    rec.update({"read_s": np.array(s_list, np.int32), "read_minhits": np.array(mh_list, np.int32), "cand_off": np.array(off, np.int64),
                "q_off": np.array(q_off, np.int64), "q_hash": np.concatenate(q_hash), "q_strand": np.concatenate(q_strand)})
    for kk in keys:
        rec["cand_" + kk] = np.concatenate(acc[kk]) if acc[kk] else np.zeros(0, np.int32)
    # minimizer streams of a few awkward sequences straight from CommonFunc::addMinimizers
    odd = [b"ACGT" * 10, b"A" * 500, b"ACACACACAC" * 60,
           b"acgtnnnnnnnnnnnnnnnnnnnnacgtacgatcgatcgatgctagctagctagctagcatgcatgcatgcatgcNNNNNNNNNNNNNNNNNNNNNNNNNNNNACGATCGATCGATCGACTGATCGATCG",
           bytes(np.random.default_rng(1).choice(list(b"ACGTacgtNRY"), size=5000,
                                                  p=[.22, .22, .22, .22, .02, .02, .02, .02, .02, .01, .01]).astype(np.uint8))]
    for i, sq in enumerate(odd):
        for (kk, ww) in ((16, 13), (16, 1), (11, 5), (8, 40), (5, 200)):
            hh, wp, st = R.minimizers(sq, kk, ww)
            rec[f"odd{i}_k{kk}_w{ww}_hash"] = hh; rec[f"odd{i}_k{kk}_w{ww}_wpos"] = wp; rec[f"odd{i}_k{kk}_w{ww}_strand"] = st
    rec["odd_seqs"] = np.array([s.decode() for s in odd])
    np.savez_compressed(os.path.join(HERE, "ref_small_internals.npz"), **rec)
    shutil.rmtree(tmp)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
