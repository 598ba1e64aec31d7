"""The oracle (oracle/mm_oracle.cpp) pinned against the reference: its own example output, outputs of the
unmodified reference binary on the `small` workload, and SciPy's Boost-backed binomial functions."""
import json
import os
import re

import numpy as np
import pytest

from metamaps_b200 import synth
from tests.conftest import GOLDEN
from tests.common import ODD_PARAMS


def g6(x):
    return "%g" % x


def test_minimizer_streams_match_reference(oracle, golden):
    odd = [s.encode() for s in golden["odd_seqs"]]
    for i, s in enumerate(odd):
        for (k, w) in ODD_PARAMS:
            h, wp, st = oracle.minimizers(s, k, w)
            assert np.array_equal(h, golden[f"odd{i}_k{k}_w{w}_hash"])
            assert np.array_equal(wp, golden[f"odd{i}_k{k}_w{w}_wpos"])
            assert np.array_equal(st, golden[f"odd{i}_k{k}_w{w}_strand"])


def test_index_and_mapping_internals_match_reference(oracle, golden, small_workload):
    k, w = int(golden["k"]), int(golden["w"])
    contigs = [synth.codes_to_ascii(c) for c in small_workload["db"].contig_codes]
    h = oracle.index_build(contigs, k, w)
    hs, sq, wp, st = oracle.index_get(h)
    assert np.array_equal(hs, golden["index_hash"]) and np.array_equal(sq, golden["index_seq"])
    assert np.array_equal(wp, golden["index_wpos"]) and np.array_equal(st, golden["index_strand"])
    assert oracle.index_unique(h) == int(golden["index_unique"])
    assert oracle.index_freq_threshold(h) == int(golden["freq_threshold"])
    off = golden["cand_off"]; qoff = golden["q_off"]
    for i, rd in enumerate(small_workload["reads"]):
        s = synth.codes_to_ascii(rd)
        if len(s) < 1000:
            assert off[i] == off[i + 1]
            continue
        m = oracle.map_read(h, s)
        assert m["s"] == golden["read_s"][i] and m["minimumHits"] == golden["read_minhits"][i]
        for key in ("seq", "start", "end", "shared", "votes", "valid", "optStart", "optEnd"):
            assert np.array_equal(m[key], golden["cand_" + key][off[i]:off[i + 1]]), (i, key)
        qh, qw, qs = oracle.read_sketch(s, k, w)
        assert np.array_equal(qh, golden["q_hash"][qoff[i]:qoff[i + 1]])
        assert np.array_equal(qs, golden["q_strand"][qoff[i]:qoff[i + 1]])
    oracle.index_free(h)


def test_example_output_identity_kat(oracle):
    """All 2985 mappings of the reference's own example run: col 10 and col 13 follow from col 11/12."""
    n = 0
    for line in open(os.path.join(GOLDEN, "example_kat.tsv")):
        if line.startswith("#"):
            continue
        rl, c10, c11, c12, c13 = line.rstrip("\n").split("\t")
        nuc, up = oracle.identity(int(c11), int(c12), 16)
        assert g6(nuc) == c10
        corrected = np.float32(np.exp(-(1 - float(c10) / 100.0)))
        assert g6(np.float32(corrected * np.float32(100))) == c13
        n += 1
    assert n == 2985


def test_binomial_matches_boost_vectors(oracle):
    d = json.load(open(os.path.join(GOLDEN, "boost_binomial.json")))
    for c in d["cases"]:
        pmf = oracle.binom_pmf(c["k"], c["n"], c["p"])
        assert abs(pmf - c["pmf"]) <= 2e-11 * max(c["pmf"], 1e-300) + 1e-300, c
        # SciPy rounds the upper quantile up exactly like Boost's default policy does for a complement
        assert oracle.binom_quantile_upper(c["n"], c["p"], c["q"]) == c["isf"], c
        sf = oracle.binom_sf(c["k"], c["n"], c["p"])
        assert abs(sf - c["sf"]) <= 1e-10 * max(c["sf"], 1e-30) + 1e-15, c


def _parse_mappings(path):
    reads = {}
    order = []
    for line in open(path):
        f = line.rstrip("\n").split(" ")
        if f[0] not in reads:
            reads[f[0]] = []
            order.append(f[0])
        reads[f[0]].append(f)
    return order, reads


def test_mapq_matches_reference_binary(oracle):
    order, reads = _parse_mappings(os.path.join(GOLDEN, "ref_small", "ref"))
    for name in order:
        rows = reads[name]
        ident = np.array([float(r[9]) / 100.0 for r in rows])
        sh = np.array([int(r[10]) for r in rows], np.int32); sk = np.array([int(r[11]) for r in rows], np.int32)
        rc, q = oracle.mapq(ident, sh, sk, int(rows[0][1]), 16)
        assert rc == 0
        for r, qq in zip(rows, q):
            assert g6(qq) == r[13] or abs(qq - float(r[13])) <= 2e-6 * max(qq, 1e-300)
            assert g6(np.float32(np.float32(np.exp(-(1 - float(r[9]) / 100.0))) * np.float32(100))) == r[12]
            nuc, up = oracle.identity(int(r[10]), int(r[11]), 16)
            assert g6(nuc) == r[9] and up >= 80.0


def em_inputs_from_files(mapping_path, taxon_info_path):
    """fEM.h:234-348 on the text files: taxon index, mapq, nloc per mapping, read offsets."""
    contig_len = {}; taxon_contigs = {}
    for line in open(taxon_info_path):
        t, rest = line.rstrip("\n").split(" ")
        for c in rest.split(";"):
            name, ln = c.rsplit("=", 1)
            contig_len[name] = int(ln); taxon_contigs.setdefault(t, []).append(name)
    order, reads = _parse_mappings(mapping_path)
    taxa = sorted({re.search(r"kraken:taxid\|(x?\d+)", r[5]).group(1) for rows in reads.values() for r in rows})
    tidx = {t: i for i, t in enumerate(taxa)}
    tax, mq, nloc, off = [], [], [], [0]
    for name in order:
        rows = reads[name]; L = int(rows[0][1])
        seen = {r[5] for r in rows}
        per_t = {}
        for r in rows:
            t = re.search(r"kraken:taxid\|(x?\d+)", r[5]).group(1)
            if t not in per_t:
                n = 0
                for c in taxon_contigs[t]:
                    if contig_len[c] >= L:
                        n += contig_len[c] - L + 1
                    elif c in seen:
                        n += 1
                per_t[t] = n
            tax.append(tidx[t]); mq.append(float(r[13])); nloc.append(per_t[t])
        off.append(len(tax))
    return taxa, order, reads, np.array(tax, np.int32), np.array(mq), np.array(nloc, np.float64), np.array(off, np.int64)


def test_em_matches_reference_binary(oracle, small_workload):
    ref = os.path.join(GOLDEN, "ref_small")
    taxa, order, reads, tax, mq, nloc, off = em_inputs_from_files(os.path.join(ref, "ref"),
                                                                 os.path.join(small_workload["dir"], "db", "taxonInfo.txt"))
    res = oracle.em(tax, mq, nloc, off, len(taxa))
    n_rounds = len(re.findall(r"^EM round", open(os.path.join(ref, "classify.log")).read(), re.M))
    assert res["iters"] == n_rounds
    lls = [float(x) for x in re.findall(r"Log likelihood: (\S+)", open(os.path.join(ref, "classify.log")).read())]
    for a, b in zip(res["ll"], lls):
        assert g6(a) == g6(b)
    # posterior column of the .EM file (std::to_string -> 6 decimals) and the read -> taxon calls
    em_lines = open(os.path.join(ref, "ref.EM")).read().splitlines()
    for m, line in enumerate(em_lines):
        assert abs(float(line.split(" ")[13]) - res["posterior"][m]) <= 5.1e-7
    r2t = dict(l.rstrip("\n").split("\t") for l in open(os.path.join(ref, "ref.EM.reads2Taxon")))
    for r, name in enumerate(order):
        assert r2t[name] == taxa[tax[res["best"][r]]]
    # EM frequencies as printed in the WIMP file (definedGenomes level, EMFrequency column, 6 significant digits)
    nz = 0
    for line in open(os.path.join(ref, "ref.EM.WIMP")):
        c = line.rstrip("\n").split("\t")
        if c[0] == "definedGenomes" and c[1] in taxa:
            assert abs(float(c[4]) - res["f"][taxa.index(c[1])]) <= 1e-6
            nz += 1
    assert nz > 0


def test_oracle_matches_live_reference_on_random_sequences(oracle, ref_harness):
    rng = np.random.default_rng(2024)
    for trial in range(40):
        n = int(rng.integers(20, 3000))
        alphabet = [b"ACGT", b"ACGTN", b"AC", b"ACGTacgtn"][trial % 4]
        s = bytes(rng.choice(list(alphabet), size=n).astype(np.uint8))
        k = int(rng.integers(4, 17)); w = int(rng.integers(1, 60))
        a = oracle.minimizers(s, k, w); b = ref_harness.minimizers(s, k, w)
        for x, y in zip(a, b):
            assert np.array_equal(x, y), (trial, k, w)
    for s_ in (1, 2, 17, 50, 300, 723, 1500, 4000):
        assert oracle.min_hits_relaxed(s_, 16, 80.0) == ref_harness.min_hits_relaxed(s_, 16, 80.0)
        for sh in (0, 1, s_ // 20, s_ // 5, s_):
            assert oracle.identity(sh, s_, 16) == ref_harness.identity(sh, s_, 16)
    for (m, L) in ((1000, 2_025_370), (2000, 12_000_000_000), (1000, 26_762_276_280), (5000, 486_296)):
        assert oracle.recommended_window(1e-3, 16, 80.0, m, L) == ref_harness.recommended_window(1e-3, 16, 80.0, m, L)
