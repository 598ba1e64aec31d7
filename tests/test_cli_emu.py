"""The C++ host linked against the host-emulation library writes the same files as the reference CLI."""
import os
import shutil
import subprocess

import pytest

from metamaps_b200 import synth
from tests import cli_common
from tests.conftest import GOLDEN, build_emu_host


def test_cli_matches_golden_reference_files(small_workload):
    binary = build_emu_host()
    got = cli_common.run_cli(binary, small_workload["dir"], out="out_emu")
    n_identical = cli_common.compare_dirs(os.path.join(GOLDEN, "ref_small"), got)
    assert n_identical >= 8      # byte-identical in practice; float columns may differ in the last printed digit


def test_cli_matches_live_reference_with_options(tmp_path):
    from oracle import pyoracle
    if not pyoracle.ref_available():
        pytest.skip("oracle/_ref not built")
    binary = build_emu_host()
    d = str(tmp_path)
    db = synth.make_db(5, 3, 2, 50_000, 0.03)
    synth.write_db(db, os.path.join(d, "db"))
    names, reads, _ = synth.make_reads(db, 6, 150, 2500, lognormal_sigma=0.4, clip=(600, 9000), frac_random=0.03)
    synth.write_fastq(os.path.join(d, "reads.fq"), names, reads)
    for extra, tag in ((("-m", "800", "--pi", "85"), "a"), (("-w", "7", "-k", "15"), "b")):
        os.makedirs(os.path.join(d, "out_ref" + tag), exist_ok=True)
        subprocess.run([pyoracle.REF_BIN, "mapDirectly", "--all", "-r", "db/DB.fa", "-q", "reads.fq", "-o", f"out_ref{tag}/ref", *extra], cwd=d, check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        subprocess.run([pyoracle.REF_BIN, "classify", "--DB", "db", "--mappings", f"out_ref{tag}/ref"], cwd=d, check=True, stdout=subprocess.DEVNULL,
                       stderr=subprocess.DEVNULL)
        got = cli_common.run_cli(binary, d, extra, out="out_emu" + tag)
        cli_common.compare_dirs(os.path.join(d, "out_ref" + tag), got)


def test_cli_without_all_keeps_top_mappings(tmp_path):
    """Default (no --all): only mappings within 1 % identity of the best are reported (computeMap.hpp:561-563)."""
    from oracle import pyoracle
    if not pyoracle.ref_available():
        pytest.skip("oracle/_ref not built")
    binary = build_emu_host()
    d = str(tmp_path)
    db = synth.make_db(9, 2, 3, 40_000, 0.02)
    synth.write_db(db, os.path.join(d, "db"))
    names, reads, _ = synth.make_reads(db, 10, 80, 3000)
    synth.write_fastq(os.path.join(d, "reads.fq"), names, reads)
    os.makedirs(os.path.join(d, "o1")); os.makedirs(os.path.join(d, "o2"))
    subprocess.run([pyoracle.REF_BIN, "mapDirectly", "-r", "db/DB.fa", "-q", "reads.fq", "-o", "o1/ref"], cwd=d, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    subprocess.run([binary, "mapDirectly", "-r", "db/DB.fa", "-q", "reads.fq", "-o", "o2/ref"], cwd=d, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    assert open(os.path.join(d, "o1", "ref")).read() == open(os.path.join(d, "o2", "ref")).read()
    assert open(os.path.join(d, "o1", "ref.meta")).read() == open(os.path.join(d, "o2", "ref.meta")).read()


def test_cli_index_then_mapAgainstIndex(small_workload):
    """`index` + `mapAgainstIndex` (GPU-native index file behind the reference's <prefix>.index manifest) = `mapDirectly`."""
    binary = build_emu_host()
    direct = cli_common.run_cli(binary, small_workload["dir"], out="out_emu_d")
    via = cli_common.run_cli_via_index(binary, small_workload["dir"], out="out_emu_ix")
    cli_common.compare_mapping_files(direct, via)
    assert cli_common.compare_dirs(os.path.join(GOLDEN, "ref_small"), via) >= 8


def test_cli_db_with_N_runs_lowercase_and_iupac(tmp_path):
    cli_common.check_db_with_N_runs_lowercase_and_iupac(build_emu_host(), tmp_path)


def test_cli_chunked_reference_equals_single_index(small_workload):
    binary = build_emu_host()
    assert cli_common.check_chunked_equals_direct(binary, small_workload["dir"]) >= 3


def test_cli_config2_shaped_sample_matches_reference_files(tmp_path):
    """408 Mbp of the config-2 recipe: the reference's occurrence threshold is FINITE there (>= 12, its own INFO line), the
    32-bit hash space starts to saturate, and all ten files must equal what the unmodified reference wrote (fixture)."""
    identical, thr = cli_common.check_config2_golden(build_emu_host(), str(tmp_path))
    assert identical == 10, identical


def test_cli_awkward_fastx_formatting_matches_live_reference(tmp_path, monkeypatch):
    """kseq record semantics (kseq.h:171-208) through the block-wise parser: wrapped FASTQ/FASTA lines, CRLF, blank lines,
    quality strings containing '@' '>' '+', comments after the name, a '+name' separator, lower case, no trailing newline;
    read batches of ~20 kB so that the parser / GPU / writer hand-over happens many times."""
    from oracle import pyoracle
    if not pyoracle.ref_available():
        pytest.skip("oracle/_ref not built")
    import numpy as np
    binary = build_emu_host()
    d = str(tmp_path)
    db = synth.make_db(31, 3, 2, 40_000, 0.02)
    synth.write_db(db, os.path.join(d, "db"), line_width=61)
    names, reads, _ = synth.make_reads(db, 32, 90, 2600, lognormal_sigma=0.3, clip=(700, 6000), frac_short=0.05, short_len=500)
    rng = np.random.default_rng(4)
    with open(os.path.join(d, "reads.fq"), "wb") as f:
        for i, (n, r) in enumerate(zip(names, reads)):
            s = synth.codes_to_ascii(r)
            if i % 5 == 1:
                s = s.lower()
            q = bytes(rng.choice(list(b"@>+I5#~!"), size=len(s)).astype(np.uint8))
            eol = b"\r\n" if i % 4 == 2 else b"\n"
            if i % 3 == 0:          # wrapped sequence and quality
                wrap = lambda x: eol.join(x[j:j + 70] for j in range(0, len(x), 70))
                f.write(b"@" + n.encode() + b" some comment" + eol + wrap(s) + eol + b"+" + n.encode() + eol + wrap(q) + eol)
            elif i % 7 == 3:        # a FASTA record in the middle of the FASTQ
                f.write(b">" + n.encode() + b"\tx=1" + eol + s + eol + eol)
            else:
                f.write(b"@" + n.encode() + eol + s + eol + b"+" + eol + q + (b"" if i == len(names) - 1 else eol))
    for out, b in (("o_ref", pyoracle.REF_BIN), ("o_emu", binary)):
        os.makedirs(os.path.join(d, out))
        if b == binary:
            monkeypatch.setenv("MM_HOST_BATCH_BYTES", "20000")
        subprocess.run([b, "mapDirectly", "--all", "-r", "db/DB.fa", "-q", "reads.fq", "-o", out + "/ref", "-t", "3"], cwd=d, check=True, stdout=subprocess.DEVNULL,
                       stderr=subprocess.DEVNULL)
        subprocess.run([b, "classify", "--DB", "db", "--mappings", out + "/ref"], cwd=d, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    assert cli_common.compare_dirs(os.path.join(d, "o_ref"), os.path.join(d, "o_emu")) >= 8
    assert int(open(os.path.join(d, "o_emu", "ref.meta")).read().split()[1]) == len(names)


def test_cli_maxmemory_chunk_loop_matches_reference_fixture(tmp_path):
    thr = cli_common.check_maxmemory_golden(build_emu_host(), str(tmp_path))
    assert [t.split(">= ")[1].split()[0] for t in thr] == ["4", "7", "23"]
