"""Named parity cases: small seeded inputs the oracle finishes in well under a second.

Each case is (fasta text as uint8 array, templates per category, simulate_batch kwargs).
The same builders feed the oracle, the unmodified reference (tests/golden/make_golden.py)
and the CUDA path, so all three see identical inputs.
"""
from __future__ import annotations

import numpy as np

from qtip_b200 import synth
from qtip_b200.templates import TemplatePaired, TemplateUnpaired


def cli_seed(seed: int):
    """qtip_parse.cpp:1669-1674"""
    return (seed, (seed * 77) % 2147483562 + 1)


def case_c1_lambda():
    """BASELINE config 1: 48.5 kbp single record, 10k unpaired 100 bp templates with 0-3 X,
    default sqrt/45/min 30000, seed 99099."""
    fa = synth.make_fasta([("lambda_like", 48502)], seed=99099)
    tu = synth.templates_unpaired(10000, 100, seed=99099, max_x=3)
    return dict(fasta=fa, tu=tu, tb=[], tc=[], td=[],
                kw=dict(fraction=45.0, function=2, min_u=30000, min_c=30000, min_d=10000, min_b=10000,
                        cg=cli_seed(99099)))


def case_mixed_small():
    """All four categories, indel templates, leading N runs, lower-case bases, a record
    shorter than olap, a record just over one window."""
    fa = synth.make_fasta([("chrA desc", 300000), ("chrB", 150001), ("tiny", 120), ("chrC", 131072 + 77)], seed=5,
                          lead_n=2000, lowercase_frac=0.1)
    tu = synth.templates_unpaired(500, 100, 1, max_x=3) + synth.templates_unpaired(100, 200, 2, indel=True)
    tb = synth.templates_badend(200, 100, 3, opp_len=50)
    tc = synth.templates_paired(300, 100, 4, frag_mean=300, frag_sd=30)
    td = synth.templates_paired(200, 100, 5, discordant=True, max_fraglen=5000)
    return dict(fasta=fa, tu=tu, tb=tb, tc=tc, td=td,
                kw=dict(fraction=3000.0, function=3, min_u=1, min_c=1, min_d=1, min_b=1, cg=(34, 2619)))


def case_c2_pairs():
    """BASELINE config 2 shape, scaled: concordant 2x150 pairs, fraglen N(400,40), 0-4 X."""
    fa = synth.make_fasta([("ecoli_like", 460000)], seed=21)
    tc = synth.templates_paired(3000, 150, 22, frag_mean=400, frag_sd=40, max_x=4)
    return dict(fasta=fa, tu=[], tb=[], tc=tc, td=[],
                kw=dict(fraction=10000.0, function=3, min_u=1, min_c=1, min_d=1, min_b=1, cg=cli_seed(2222)))


def case_c5_indel():
    """BASELINE config 5 shape, scaled: 1 kbp reads with ~10% edited columns, indel runs 1-6."""
    fa = synth.make_fasta([("ecoli_like", 460000)], seed=51)
    tu = synth.templates_unpaired(1000, 1000, 52, indel=True)
    return dict(fasta=fa, tu=tu, tb=[], tc=[], td=[],
                kw=dict(fraction=2000.0, function=3, min_u=1, min_c=1, min_d=1, min_b=1, cg=cli_seed(555)))


def case_c3_shape_small():
    """BASELINE config 3 shape, scaled 1/2000: 24 GRCh38-proportioned records with leading
    N runs, unpaired 250 bp templates with 0-5 X."""
    fa = synth.make_fasta(synth.grch38_like_records(1.0 / 2000), seed=31, lead_n=500)
    tu = synth.templates_unpaired(2000, 250, 32, max_x=5)
    return dict(fasta=fa, tu=tu, tb=[], tc=[], td=[],
                kw=dict(fraction=5000.0, function=3, min_u=1, min_c=1, min_d=1, min_b=1, cg=cli_seed(333)))


def case_c4_mix_small():
    """BASELINE config 4 shape, scaled: c/d/b templates 2x150, 70/15/15 mix."""
    fa = synth.make_fasta(synth.grch38_like_records(1.0 / 4000), seed=41, lead_n=300)
    tc = synth.templates_paired(1500, 150, 42, frag_mean=400, frag_sd=40, max_x=4)
    td = synth.templates_paired(1500, 150, 43, discordant=True, max_fraglen=20000)
    tb = synth.templates_badend(1500, 150, 44, opp_len=150)
    return dict(fasta=fa, tu=[], tb=tb, tc=tc, td=td, added=(0, 1500, 7000, 1500),
                kw=dict(fraction=1.0, function=1, min_u=1, min_c=1, min_d=1, min_b=1, cg=cli_seed(444)))


def case_edges():
    """Window edge cases with a small chunk size: records of exactly chunksz, chunksz+1,
    olap-1, olap and olap+1 bases; a >90%-N window; IUPAC codes; CRLF; blank lines; a
    header with a tab; bases before the first header."""
    rng = np.random.default_rng(77)
    chunksz, L = 4096, 100

    def rec(name, n, width=60, nl=b"\n"):
        s = synth.random_bases(n, rng).tobytes()
        lines = [s[i:i + width] for i in range(0, len(s), width)]
        return b">" + name + nl + nl.join(lines) + nl

    parts = [b"ACGTACGTNNACGTTTGA" * 20 + b"\n"]  # headerless leading bases -> record ""
    parts.append(rec(b"exact", chunksz))
    parts.append(rec(b"plus1 some description", chunksz + 1))
    parts.append(rec(b"olapm1", L - 1))
    parts.append(rec(b"olap", L))
    parts.append(rec(b"olapp1", L + 1))
    parts.append(rec(b"crlf\tx y", 3 * chunksz + 17, width=70, nl=b"\r\n"))
    parts.append(b">mostlyN\n" + b"N" * 3900 + synth.random_bases(196, rng).tobytes() + b"\n" +
                 synth.random_bases(5000, rng).tobytes() + b"\n")
    parts.append(b">iupac\n" + b"ACGTRYKMSWBDHVNacgtn-ACGT" * 400 + b"\n\n\n")
    parts.append(rec(b"last_no_newline", 9000)[:-1])
    fa = np.frombuffer(b"".join(parts), dtype=np.uint8).copy()
    tu = synth.templates_unpaired(300, L, 78, max_x=6) + synth.templates_unpaired(50, L, 79, indel=True, len_jitter=30)
    # trim indel templates whose reference span exceeds L so olap stays L
    tu = [t for t in tu if t.reflen() <= L]
    return dict(fasta=fa, tu=tu, tb=[], tc=[], td=[], chunksz=chunksz,
                kw=dict(fraction=4000.0, function=3, min_u=1, min_c=1, min_d=1, min_b=1, cg=cli_seed(7)))


def case_inversion():
    """Small targets: every ignbin call takes the inversion branch (n*p < 30,
    ranlib.cpp:1198-1231), including categories with n = 0."""
    fa = synth.make_fasta([("a", 700000), ("b", 90000)], seed=61)
    tu = synth.templates_unpaired(40, 75, 62, max_x=2)
    tc = synth.templates_paired(30, 75, 63, frag_mean=250, frag_sd=20)
    return dict(fasta=fa, tu=tu, tb=[], tc=tc, td=[],
                kw=dict(fraction=60.0, function=3, min_u=1, min_c=1, min_d=1, min_b=1, cg=cli_seed(66)))


def case_long_refid():
    """A refid longer than the kernel's shared-memory name staging, all pair types."""
    name = "contig_" + "x" * 400
    fa = synth.make_fasta([(name + " description", 200000)], seed=71)
    tb = synth.templates_badend(100, 80, 72, opp_len=40)
    tc = synth.templates_paired(100, 80, 73, frag_mean=300, frag_sd=30)
    tu = synth.templates_unpaired(100, 80, 74)
    return dict(fasta=fa, tu=tu, tb=tb, tc=tc, td=[],
                kw=dict(fraction=300.0, function=3, min_u=1, min_c=1, min_d=1, min_b=1, cg=cli_seed(71)))


def case_refid_lengths():
    """Refids of 1 .. 47 characters (the simulate kernels print refids of up to 16 bytes from registers and longer
    ones byte by byte; names decide how many edge chunks a record has), all categories, offsets of one to six digits."""
    names = ["a", "chr7xyz", "chr8abcd", "contig_09", "NC_000001.11_xx", "NC_000001.11_xyz", "gi|568336023|gb|A",
             "gi|568336023|gb|CM000663", "scaffold_0123456789_0123456789_01",
             "very_long_contig_name_0123456789_0123456789_012"]
    assert sorted(len(n) for n in names) == [1, 7, 8, 9, 15, 16, 17, 24, 33, 47]
    fa = synth.make_fasta([(n + " some description", 9000 + 7000 * k) for k, n in enumerate(names)], seed=91, lead_n=30)
    tu = synth.templates_unpaired(80, 100, 92, max_x=3)
    tb = synth.templates_badend(60, 75, 93, opp_len=60)
    tc = synth.templates_paired(80, 100, 94, frag_mean=320, frag_sd=30, max_x=3)
    td = synth.templates_paired(40, 60, 95, discordant=True, max_fraglen=2500)
    return dict(fasta=fa, tu=tu, tb=tb, tc=tc, td=td,
                kw=dict(fraction=1500.0, function=3, min_u=1, min_c=1, min_d=1, min_b=1, cg=cli_seed(91)))


def case_short_reads():
    """Bodies shorter than one 16-byte output chunk next to normal ones, all categories: the direct
    kernel declines the category (first / last chunk of a string could coincide) and the staged kernel
    takes it."""
    fa = synth.make_fasta([("s1", 90000), ("s2 x", 40000)], seed=81, lead_n=100)
    tu = (synth.templates_unpaired(60, 12, 82, max_x=2) + synth.templates_unpaired(60, 17, 83, max_x=2) +
          synth.templates_unpaired(60, 40, 84, indel=True) + synth.templates_unpaired(40, 5, 85, max_x=1))
    tb = synth.templates_badend(80, 30, 86, opp_len=8)
    tc = synth.templates_paired(80, 14, 87, frag_mean=60, frag_sd=8, max_x=2)
    td = synth.templates_paired(60, 33, 88, discordant=True, max_fraglen=3000)
    return dict(fasta=fa, tu=tu, tb=tb, tc=tc, td=td,
                kw=dict(fraction=400.0, function=3, min_u=1, min_c=1, min_d=1, min_b=1, cg=cli_seed(81)))


def case_long_reads():
    """20 kbp reads (1250 output chunks per string): substitution-only templates on the fast path and
    indel-heavy ones on the general path, both strands; tiles shrink to a few reads per warp."""
    fa = synth.make_fasta([("ctgL", 400000), ("ctgM", 150000)], seed=91, lead_n=500)
    tu = synth.templates_unpaired(12, 20000, 92, max_x=40) + synth.templates_unpaired(8, 20000, 93, indel=True)
    tc = synth.templates_paired(6, 9000, 94, frag_mean=25000, frag_sd=500, max_x=20)
    return dict(fasta=fa, tu=tu, tb=[], tc=tc, td=[],
                kw=dict(fraction=60.0, function=3, min_u=1, min_c=1, min_d=1, min_b=1, cg=cli_seed(91)))


CASES = {
    "c1_lambda": case_c1_lambda,
    "mixed_small": case_mixed_small,
    "c2_pairs": case_c2_pairs,
    "c3_shape_small": case_c3_shape_small,
    "c4_mix_small": case_c4_mix_small,
    "c5_indel": case_c5_indel,
    "edges": case_edges,
    "inversion": case_inversion,
    "long_refid": case_long_refid,
    "short_reads": case_short_reads,
    "refid_lengths": case_refid_lengths,
    "long_reads": case_long_reads,
}


# ---- BASELINE.json configurations at their stated sizes (tests/golden/make_golden_full.py runs the unmodified
# reference on them; tests/test_gpu_fullsize.py compares the CUDA path's streams with those digests) ----

def full_c2():
    """BASELINE config 2 at full size: 4.6 Mbp single record, 30 k concordant 2x150 templates (fraglen N(400,40),
    0-4 X), const 1 M pairs."""
    fa = synth.make_fasta([("ecoli_like", 4_600_000)], seed=2021)
    tc = synth.templates_paired(30000, 150, 2022, frag_mean=400, frag_sd=40, max_x=4)
    return dict(fasta=fa, tu=[], tb=[], tc=tc, td=[],
                kw=dict(fraction=1_000_000.0, function=3, min_u=1, min_c=1, min_d=1, min_b=1, cg=cli_seed(20202)))


def full_c5():
    """BASELINE config 5 at full size: 4.6 Mbp genome, 10 k unpaired 1 kbp indel-heavy templates, const 200 k."""
    fa = synth.make_fasta([("ecoli_like", 4_600_000)], seed=5051)
    tu = synth.templates_unpaired(10000, 1000, 5052, indel=True)
    return dict(fasta=fa, tu=tu, tb=[], tc=[], td=[],
                kw=dict(fraction=200_000.0, function=3, min_u=1, min_c=1, min_d=1, min_b=1, cg=cli_seed(50505)))


def _from_bench(which, n_gpus, scale, seed):
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    import bench
    wl = bench.workload(n_gpus, scale, which)
    fa = synth.make_fasta(wl["records"], seed=wl["genome_seed"], lead_n=wl["lead_n"])
    case = dict(fasta=fa, tu=wl["tu"], tb=wl["tb"], tc=wl["tc"], td=wl["td"], kw=dict(wl["kw"], cg=cli_seed(seed)))
    if wl.get("added"):
        case["added"] = wl["added"]
    return case


def full_c3():
    """BASELINE config 3 at full size = the benchmark's N=1 workload: 3.1 Gbp in 24 records, 30 k unpaired 250 bp
    templates (0-5 X), const 10 M."""
    return _from_bench("c3", 1, 1.0, 30303)


def full_c4_tenth():
    """BASELINE config 4 at 1/10: 0.31 Gbp, 10 M pairs, 70/15/15 concordant / discordant / bad-end 2x150 -- the
    benchmark's N=8 workload with genome and targets scaled by 0.1."""
    return _from_bench("c4", 8, 0.1, 40404)


FULL_CASES = {"c2_full": full_c2, "c5_full": full_c5, "c3_full": full_c3, "c4_tenth": full_c4_tenth}


def random_case(seed: int) -> dict:
    """A small configuration drawn from the seed: record layout (incl. a record below the overlap), line width,
    lower-case fraction, window size (down to longest template + 1), target function with `added` counts, read
    length, indel / soft-clipped transcripts and the category mix."""
    import random
    rnd = random.Random(seed)
    L = rnd.choice([30, 75, 100, 151])
    recs = [(f"r{k}", rnd.randint(200, 60000)) for k in range(rnd.randint(1, 5))]
    recs.insert(rnd.randint(0, len(recs)), ("short", rnd.randint(1, L)))
    fa = synth.make_fasta(recs, seed=seed, lead_n=rnd.choice([0, 0, 50, 700]), width=rnd.choice([50, 60, 70, 80]),
                          lowercase_frac=rnd.choice([0.0, 0.2]))
    indel = rnd.random() < 0.5
    tu = synth.templates_unpaired(rnd.randint(0, 40), L, seed, max_x=4, indel=indel, len_jitter=rnd.choice([0, 10]))
    tb = synth.templates_badend(rnd.randint(0, 20), L, seed + 1, opp_len=rnd.choice([20, L])) if rnd.random() < 0.7 else []
    tc = synth.templates_paired(rnd.randint(0, 30), L, seed + 2, frag_mean=3 * L, frag_sd=L / 4, indel=indel) if rnd.random() < 0.7 else []
    td = synth.templates_paired(rnd.randint(0, 20), L, seed + 3, discordant=True, max_fraglen=2000) if rnd.random() < 0.7 else []
    if not (tu or tb or tc or td):
        tu = synth.templates_unpaired(5, L, seed, max_x=2)
    longest = max([t.reflen() for t in tu + tb] + [t.fraglen for t in tc + td])
    chunksz = rnd.choice([longest + 1, longest + 17, 2 * longest + 3, 4096 + longest, 128 * 1024])
    function = rnd.choice([1, 2, 3])
    fraction = {1: rnd.choice([0.5, 2.0, 7.5]), 2: rnd.choice([3.0, 45.0]), 3: rnd.choice([1.0, 200.0, 1500.0])}[function]
    added = (len(tu) * rnd.randint(1, 30), len(tb) * rnd.randint(1, 30), len(tc) * rnd.randint(1, 30), len(td) * rnd.randint(1, 30))
    return dict(fasta=fa, tu=tu, tb=tb, tc=tc, td=td, chunksz=chunksz, added=added,
                kw=dict(fraction=fraction, function=function, min_u=rnd.choice([0, 1, 300]), min_c=rnd.choice([0, 1, 300]),
                        min_d=rnd.choice([0, 1, 100]), min_b=rnd.choice([0, 1, 100]), cg=cli_seed(seed)))


RANDOM_SEEDS = list(range(100, 148))


def sim_kwargs(case: dict) -> dict:
    kw = dict(case["kw"])
    kw["chunksz"] = case.get("chunksz", 128 * 1024)
    kw["added"] = case.get("added")
    return kw
