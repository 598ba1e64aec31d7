"""File-level parity of the C++ host (`metamaps mapDirectly` + `classify`) with the reference's files."""
import os
import subprocess

OUTPUTS = ["ref", "ref.meta", "ref.meta.unmappedReadsLengths", "ref.EM", "ref.EM.WIMP", "ref.EM.reads2Taxon", "ref.EM.reads2Taxon.krona",
           "ref.EM.contigCoverage", "ref.EM.lengthAndIdentitiesPerMappingUnit", "ref.EM.evidenceUnknownSpecies"]
FLOAT_COLS = {"ref": (" ", {9, 12, 13}), "ref.EM": (" ", {9, 12, 13}), "ref.EM.WIMP": ("\t", {4, 5}), "ref.EM.reads2Taxon.krona": ("\t", {2}),
              "ref.EM.contigCoverage": ("\t", {6}), "ref.EM.lengthAndIdentitiesPerMappingUnit": ("\t", {3}),
              "ref.EM.evidenceUnknownSpecies": ("\t", {4, 5, 6, 9, 11, 12})}


def run_cli(binary, workdir, extra_map=(), out="out_b200"):
    os.makedirs(os.path.join(workdir, out), exist_ok=True)
    subprocess.run([binary, "mapDirectly", "--all", "-r", "db/DB.fa", "-q", "reads.fq", "-o", f"{out}/ref", *extra_map], cwd=workdir, check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
    subprocess.run([binary, "classify", "--DB", "db", "--mappings", f"{out}/ref"], cwd=workdir, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
    return os.path.join(workdir, out)


def _isfloat(x):
    try:
        float(x)
        return True
    except ValueError:
        return False


def compare_dirs(ref_dir, got_dir):
    """Integer / text columns must be identical; float columns within 1e-6 absolute (north_star) -- in practice the
    files are byte-identical, the tolerance only absorbs last-digit rounding of 6-digit prints."""
    identical = 0
    for fn in OUTPUTS:
        a = open(os.path.join(ref_dir, fn)).read().splitlines(); b = open(os.path.join(got_dir, fn)).read().splitlines()
        assert len(a) == len(b), fn
        if a == b:
            identical += 1
            continue
        sep, cols = FLOAT_COLS.get(fn, (" ", set()))
        for la, lb in zip(a, b):
            if la == lb:
                continue
            fa, fb = la.split(sep), lb.split(sep)
            assert len(fa) == len(fb), (fn, la, lb)
            for i, (x, y) in enumerate(zip(fa, fb)):
                if x == y:
                    continue
                assert i in cols and _isfloat(x) and _isfloat(y), (fn, i, la, lb)
                assert abs(float(x) - float(y)) <= 1e-6 + 2e-6 * abs(float(x)), (fn, i, la, lb)
    pa = dict(l.split(" ", 1) for l in open(os.path.join(ref_dir, "ref.parameters")).read().splitlines())
    pb = dict(l.split(" ", 1) for l in open(os.path.join(got_dir, "ref.parameters")).read().splitlines())
    for k in pa:
        if k != "outFileName":
            assert pa[k] == pb[k], k
    return identical


def run_cli_via_index(binary, workdir, out="out_ix"):
    """`metamaps index` then `metamaps mapAgainstIndex` (+ classify): the persistent-index route to the same files."""
    os.makedirs(os.path.join(workdir, out), exist_ok=True)
    subprocess.run([binary, "index", "-r", "db/DB.fa", "-i", f"{out}/idx"], cwd=workdir, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
    manifest = open(os.path.join(workdir, out, "idx.index")).read().split()
    assert manifest[0] == "1" and len(manifest) == 2 and os.path.exists(os.path.join(workdir, manifest[1]))     # mapWrap.h:395-403
    subprocess.run([binary, "mapAgainstIndex", "--all", "-i", f"{out}/idx", "-q", "reads.fq", "-o", f"{out}/ref"], cwd=workdir, check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
    subprocess.run([binary, "classify", "--DB", "db", "--mappings", f"{out}/ref"], cwd=workdir, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
    return os.path.join(workdir, out)


def compare_mapping_files(dir_a, dir_b):
    """Every output of the index route equals the mapDirectly route byte for byte (same library, same arrays)."""
    for fn in OUTPUTS:
        assert open(os.path.join(dir_a, fn)).read() == open(os.path.join(dir_b, fn)).read(), fn


def check_db_with_N_runs_lowercase_and_iupac(binary, tmp_path):
    """A DB directory as buildDB.pl leaves it for real genomes: N runs, soft-masked lower case, IUPAC codes in
    contigs and reads, and non-zero contigNstats.  Non-ACGT bytes are hashed verbatim by the reference (commonFunc.hpp:44-51)
    -> the packed-sequence exception path of K0/K1 -- and N windows enter the coverage statistics of classify."""
    import pytest
    from metamaps_b200 import synth
    from oracle import pyoracle
    if not pyoracle.ref_available():
        pytest.skip("oracle/_ref not built")
    d = str(tmp_path)
    db = synth.make_db(9, 3, 2, 60_000, 0.02)
    synth.write_db(db, os.path.join(d, "db"))
    import numpy as np
    rng = np.random.default_rng(4)
    contigs = []
    for c in db.contig_codes:
        a = bytearray(synth.codes_to_ascii(c))
        for _ in range(3):                                   # N runs (some longer than a window of k-mers, some short)
            p = int(rng.integers(0, len(a) - 400)); n = int(rng.choice([1, 7, 40, 350]))
            a[p:p + n] = b"N" * n
        for _ in range(4):                                   # soft-masked repeats
            p = int(rng.integers(0, len(a) - 600)); a[p:p + 500] = bytes(a[p:p + 500]).lower()
        for _ in range(6):                                   # IUPAC ambiguity codes
            a[int(rng.integers(0, len(a)))] = int(rng.choice(list(b"RYKMSWn")))
        contigs.append(bytes(a))
    with open(os.path.join(d, "db", "DB.fa"), "wb") as f:
        for name, a in zip(db.contig_names, contigs):
            f.write(b">" + name.encode() + b"\n")
            for i in range(0, len(a), 80):
                f.write(a[i:i + 80] + b"\n")
    with open(os.path.join(d, "db", "contigNstats_windowSize_1000.txt"), "w") as f:
        for name, t, a in zip(db.contig_names, db.contig_taxon, contigs):
            up = a.upper()
            f.write(f"{t}\t{name}\t" + ";".join(str(up[i:i + 1000].count(b"N")) for i in range(0, len(a), 1000)) + "\n")
    names, reads, _ = synth.make_reads(db, 10, 120, 2500, lognormal_sigma=0.3, clip=(900, 6000))
    reads_asc = []
    for r in reads:
        a = bytearray(synth.codes_to_ascii(r))
        if rng.random() < 0.3:
            p = int(rng.integers(0, max(1, len(a) - 60))); a[p:p + 30] = b"N" * min(30, len(a) - p)
        if rng.random() < 0.3:
            a[:] = bytes(a).lower()
        reads_asc.append(bytes(a))
    with open(os.path.join(d, "reads.fq"), "wb") as f:
        for nme, a in zip(names, reads_asc):
            f.write(b"@" + nme.encode() + b"\n" + a + b"\n+\n" + b"I" * len(a) + b"\n")
    os.makedirs(os.path.join(d, "out_ref"), exist_ok=True)
    subprocess.run([pyoracle.REF_BIN, "mapDirectly", "--all", "-r", "db/DB.fa", "-q", "reads.fq", "-o", "out_ref/ref"], cwd=d, check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    subprocess.run([pyoracle.REF_BIN, "classify", "--DB", "db", "--mappings", "out_ref/ref"], cwd=d, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    got = run_cli(binary, d, out="out_emu")
    compare_dirs(os.path.join(d, "out_ref"), got)
    assert sum(1 for _ in open(os.path.join(d, "out_ref", "ref"))) > 100


def check_chunked_equals_direct(binary, workdir, chunk_bases=150_000):
    """`mapDirectly` with the reference cut into index chunks (the --maxmemory loop: chunk files <prefix>.N, then unifyFiles)
    writes the same files as the single-index run whenever no hash is over-frequent."""
    direct = run_cli(binary, workdir, out="out_direct")
    out = "out_chunked"
    os.makedirs(os.path.join(workdir, out), exist_ok=True)
    env = dict(os.environ, MM_HOST_CHUNK_BASES=str(chunk_bases))
    p = subprocess.run([binary, "mapDirectly", "--all", "-r", "db/DB.fa", "-q", "reads.fq", "-o", f"{out}/ref"], cwd=workdir, check=True,
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env, text=True)
    n_chunks = p.stdout.count("Index chunk ")
    assert n_chunks >= 3, p.stdout[-500:]
    assert not any(fn.startswith("ref.") and fn[4:].isdigit() for fn in os.listdir(os.path.join(workdir, out)))      # chunk files removed
    subprocess.run([binary, "classify", "--DB", "db", "--mappings", f"{out}/ref"], cwd=workdir, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
    compare_mapping_files(direct, os.path.join(workdir, out))
    return n_chunks


def check_config2_golden(binary, workdir):
    """The C++ host on the config-2-shaped sample (408 Mbp, finite occurrence threshold) against the files the UNMODIFIED
    reference wrote for it (tests/golden/ref_config2, generator make_golden_config2.py).  Returns (identical files, threshold line)."""
    import gzip
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "tests", "golden"))
    import make_golden_config2 as g
    g.config2_sample(workdir)
    gold = os.path.join(workdir, "gold"); os.makedirs(gold, exist_ok=True)
    src = os.path.join(root, "tests", "golden", "ref_config2")
    for fn in os.listdir(src):
        with gzip.open(os.path.join(src, fn), "rb") as f, open(os.path.join(gold, fn[:-3]), "wb") as o:
            o.write(f.read())
    thr = [l for l in open(os.path.join(gold, "map.log")) if "ignore minimizers occurring" in l]
    assert thr and ">= 12 times" in thr[0], "the fixture must pin a FINITE occurrence threshold"
    out = os.path.join(workdir, "out_b200"); os.makedirs(out, exist_ok=True)
    p = subprocess.run([binary, "mapDirectly", *g.MAP_ARGS, "-r", "db/DB.fa", "-q", "reads.fq", "-o", "out_b200/ref"], cwd=workdir, check=True,
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    mine = [l for l in p.stdout.splitlines() if "ignore minimizers occurring" in l]
    assert mine and mine[0].strip() == thr[0].strip(), (mine, thr)          # same computeFreqHist line as the reference
    subprocess.run([binary, "classify", "--DB", "db", "--mappings", "out_b200/ref"], cwd=workdir, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
    return compare_dirs(gold, out), thr[0].strip()


def check_maxmemory_golden(binary, workdir):
    """The reference's --maxmemory chunk loop, including the occurrence histogram / threshold that it does NOT reset between
    chunks (winSketch.hpp:302-304,452-495): our host, cutting at the same contig counts, must print the reference's three
    thresholds (4, 7, 23) and write its mapping file byte for byte (tests/golden/ref_maxmemory, make_golden_maxmemory.py)."""
    import gzip
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "tests", "golden"))
    import make_golden_maxmemory as g
    g.maxmemory_sample(workdir)
    src = os.path.join(root, "tests", "golden", "ref_maxmemory")
    log = open(os.path.join(src, "map.log")).read().splitlines()
    cuts = [l.split()[-1] for l in log if "storeCurrentState" in l]
    thr = [l.strip() for l in log if "computeFreqHist" in l]
    assert len(cuts) == 3 and len(thr) == 3 and len(set(thr)) == 3
    out = os.path.join(workdir, "o_b200"); os.makedirs(out, exist_ok=True)
    env = dict(os.environ, MM_HOST_CHUNK_CONTIGS=",".join(cuts))
    p = subprocess.run([binary, "mapDirectly", *g.MAP_ARGS, "-r", "db/DB.fa", "-q", "reads.fq", "-o", "o_b200/ref", "-t", "4"], cwd=workdir, check=True,
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env)
    mine = [l.strip() for l in p.stdout.splitlines() if "computeFreqHist" in l]
    assert mine == thr, (mine, thr)
    for fn in ("ref", "ref.meta", "ref.meta.unmappedReadsLengths"):
        want = gzip.open(os.path.join(src, fn + ".gz"), "rb").read()
        assert open(os.path.join(out, fn), "rb").read() == want, fn
    # the same chunks through the persistent-index route: `index` writes three chunk files, `mapAgainstIndex` walks them
    ix = os.path.join(workdir, "o_ix"); os.makedirs(ix, exist_ok=True)
    subprocess.run([binary, "index", "-w", "2", "-r", "db/DB.fa", "-i", "o_ix/idx"], cwd=workdir, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, env=env)
    manifest = open(os.path.join(ix, "idx.index")).read().split()
    assert manifest[0] == "1" and len(manifest) == 4, manifest                  # mapWrap.h:397-404: "1" + one line per chunk file
    subprocess.run([binary, "mapAgainstIndex", "--all", "-i", "o_ix/idx", "-q", "reads.fq", "-o", "o_ix/ref", "-t", "4"], cwd=workdir, check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
    for fn in ("ref", "ref.meta", "ref.meta.unmappedReadsLengths"):
        assert open(os.path.join(ix, fn), "rb").read() == open(os.path.join(out, fn), "rb").read(), fn
    return thr
