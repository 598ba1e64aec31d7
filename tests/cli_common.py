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
