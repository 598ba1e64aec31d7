"""Deterministic synthetic workloads: a MetaMaps DB directory and a FASTQ.

DB directory layout follows what `metamaps classify` reads
(reference: src/meta/fEM.h:1320-1358 taxonInfo.txt, src/meta/taxonomy.h:137-231 names/nodes.dmp,
src/meta/fEM.h:1421-1453 contigNstats_windowSize_1000.txt); contig names carry
`kraken:taxid|<id>` (fEM.h:1398).  Reads follow simulate.pl:57 (accuracy 0.88): independent
substitution / insertion / deletion errors at a total rate of 12 %.

Everything is ACGT-only and seeded, so the same arguments give the same bytes on every machine
with this numpy.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field

import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.array([3, 2, 1, 0], dtype=np.uint8)


@dataclass
class SynthDB:
    contig_names: list            # FASTA names, DB.fa order
    contig_taxon: list            # taxon id (str) per contig
    contig_codes: list            # np.uint8 arrays with values 0..3
    taxon_parent: dict = field(default_factory=dict)   # id -> (parent, rank, name)

    @property
    def total_bases(self) -> int:
        return int(sum(len(c) for c in self.contig_codes))


def make_db(seed: int, n_species: int, n_strains: int, contig_len: int, divergence: float,
            contigs_per_strain: int = 1, star: bool = False, div_range=None) -> SynthDB:
    """n_species ancestors; each strain = ancestor with `divergence` substitutions.

    star=True with div_range=(lo,hi): every strain draws its own divergence (config 4).
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    names, taxa, codes = [], [], []
    tax = {"1": ("1", "no rank", "root"), "2": ("1", "superkingdom", "Synthetica")}
    ci = 0
    for s in range(n_species):
        genus = str(1000 + s // 4)
        species = str(10000 + s)
        tax.setdefault(genus, ("2", "genus", f"Genus{s // 4}"))
        tax[species] = (genus, "species", f"Genus{s // 4} species{s}")
        anc = [rng.integers(0, 4, size=contig_len, dtype=np.uint8) for _ in range(contigs_per_strain)]
        for t in range(n_strains):
            strain = str(1000000 + s * 100 + t) if n_strains > 1 else species
            if n_strains > 1:
                tax[strain] = (species, "no rank", f"Genus{s // 4} species{s} strain{t}")
            d = divergence if div_range is None else float(rng.uniform(*div_range))
            for a in anc:
                c = a.copy()
                mut = rng.random(contig_len) < d
                c[mut] = (c[mut] + rng.integers(1, 4, size=int(mut.sum()), dtype=np.uint8)) & 3
                names.append(f"C{ci}|kraken:taxid|{strain}|NC_{ci:06d}.1")
                taxa.append(strain)
                codes.append(c)
                ci += 1
    return SynthDB(names, taxa, codes, tax)


def write_db(db: SynthDB, outdir: str, line_width: int = 80) -> str:
    """Write DB.fa, taxonInfo.txt, taxonomy/{nodes,names,merged}.dmp, contigNstats_windowSize_1000.txt."""
    os.makedirs(os.path.join(outdir, "taxonomy"), exist_ok=True)
    fa = os.path.join(outdir, "DB.fa")
    with open(fa, "wb") as f:
        for name, c in zip(db.contig_names, db.contig_codes):
            f.write(b">" + name.encode() + b"\n")
            asc = _ACGT[c]
            n = len(asc)
            full = (n // line_width) * line_width
            if full:
                block = np.empty((full // line_width, line_width + 1), dtype=np.uint8)
                block[:, :line_width] = asc[:full].reshape(-1, line_width)
                block[:, line_width] = 10
                f.write(block.tobytes())
            if n > full:
                f.write(asc[full:].tobytes() + b"\n")
    per_taxon: dict = {}
    for name, t, c in zip(db.contig_names, db.contig_taxon, db.contig_codes):
        per_taxon.setdefault(t, []).append((name, len(c)))
    with open(os.path.join(outdir, "taxonInfo.txt"), "w") as f:
        for t, lst in per_taxon.items():
            f.write(t + " " + ";".join(f"{n}={l}" for n, l in lst) + "\n")
    with open(os.path.join(outdir, "contigNstats_windowSize_1000.txt"), "w") as f:
        for name, t, c in zip(db.contig_names, db.contig_taxon, db.contig_codes):
            nw = max(1, (len(c) + 999) // 1000)
            f.write(f"{t}\t{name}\t" + ";".join(["0"] * nw) + "\n")
    with open(os.path.join(outdir, "taxonomy", "nodes.dmp"), "w") as fn, \
            open(os.path.join(outdir, "taxonomy", "names.dmp"), "w") as fm:
        for tid, (parent, rank, nm) in db.taxon_parent.items():
            fn.write(f"{tid}\t|\t{parent}\t|\t{rank}\t|\t\t|\n")
            fm.write(f"{tid}\t|\t{nm}\t|\t\t|\tscientific name\t|\n")
    open(os.path.join(outdir, "taxonomy", "merged.dmp"), "w").close()
    return fa


def _mutate(rng, src: np.ndarray, err: float) -> np.ndarray:
    """Independent sub/ins/del at total rate `err` (a third each)."""
    n = len(src)
    u = rng.random(n)
    e3 = err / 3.0
    out_len = np.ones(n, dtype=np.int64)
    sub = u < e3
    ins = (u >= e3) & (u < 2 * e3)
    dele = (u >= 2 * e3) & (u < 3 * e3)
    out_len[ins] = 2
    out_len[dele] = 0
    pos = np.cumsum(out_len) - out_len
    total = int(out_len.sum())
    out = np.empty(total, dtype=np.uint8)
    keep = ~dele
    b = src.copy()
    b[sub] = (b[sub] + rng.integers(1, 4, size=int(sub.sum()), dtype=np.uint8)) & 3
    out[pos[keep]] = b[keep]
    out[pos[ins] + 1] = rng.integers(0, 4, size=int(ins.sum()), dtype=np.uint8)
    return out


def make_reads(db: SynthDB, seed: int, n_reads: int, mean_len: int = 5000, lognormal_sigma: float = 0.0,
               clip=(1200, 40000), err: float = 0.12, frac_short: float = 0.0, short_len: int = 800,
               frac_random: float = 0.0, abundances=None):
    """Returns (names, list of np.uint8 code arrays, truth contig index or -1)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    ncont = len(db.contig_codes)
    if abundances is None:
        abundances = np.ones(ncont) / ncont
    names, reads, truth = [], [], []
    for r in range(n_reads):
        kind = rng.random()
        if lognormal_sigma > 0:
            mu = np.log(mean_len) - 0.5 * lognormal_sigma ** 2
            L = int(np.clip(rng.lognormal(mu, lognormal_sigma), clip[0], clip[1]))
        else:
            L = mean_len
        if kind < frac_short:
            L = short_len
        if kind >= frac_short and kind < frac_short + frac_random:
            rd = rng.integers(0, 4, size=L, dtype=np.uint8)
            ci = -1
        else:
            ci = int(rng.choice(ncont, p=abundances))
            c = db.contig_codes[ci]
            L = min(L, len(c))
            st = int(rng.integers(0, len(c) - L + 1))
            rd = _mutate(rng, c[st:st + L], err)
            if rng.random() < 0.5:
                rd = _COMP[rd[::-1]]
        names.append(f"read{r}")
        reads.append(np.ascontiguousarray(rd))
        truth.append(ci)
    return names, reads, truth


def write_fastq(path: str, names, reads) -> None:
    with open(path, "wb") as f:
        for n, r in zip(names, reads):
            f.write(b"@" + n.encode() + b"\n" + _ACGT[r].tobytes() + b"\n+\n" + b"I" * len(r) + b"\n")


def codes_to_ascii(c: np.ndarray) -> bytes:
    return _ACGT[c].tobytes()


# The five BASELINE.json configs, scaled by `scale` for tests (1.0 = as named).
def config1(outdir: str, n_reads: int = 1000, contig_len: int = 200_000):
    """config 1: seed 7, 5 species x 2 strains (2 % div.) x 200 kbp; 5 kb reads, 2 % short, 1 % random."""
    db = make_db(7, 5, 2, contig_len, 0.02)
    fa = write_db(db, os.path.join(outdir, "db"))
    names, reads, truth = make_reads(db, 8, n_reads, 5000, frac_short=0.02, frac_random=0.01)
    fq = os.path.join(outdir, "reads.fq")
    write_fastq(fq, names, reads)
    return db, fa, fq, (names, reads, truth)
