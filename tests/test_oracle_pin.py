"""Pins the CPU oracle (oracle/qtipsim_oracle.c) to the reference.

* known answers of the reference's RNG (probed from rnglib.cpp/ranlib.cpp)
* golden FASTQ digests generated from the unmodified reference (tests/golden/make_golden.py)
* the reference's own in-source unit tests: fasta.cpp:134-266 (chunk semantics) and
  simplesim.cpp:414-512 (seven read-level properties), restated against the oracle
* live comparison against oracle/_ref (the compiled reference) when it is present
"""
import ctypes as C
import os
import struct

import numpy as np
import pytest

import cases
import util
from oracle import pyoracle as po
from qtip_b200 import synth
from qtip_b200.templates import TemplateUnpaired


def f32_bits(x: float) -> int:
    return struct.unpack("<I", struct.pack("<f", x))[0]


def test_rng_known_answers():
    lib = po.orc()
    r = po.OrcRng()
    lib.orc_rng_initialize(C.byref(r))
    assert (r.cg1, r.cg2) == (1234567890, 123456789)
    assert lib.orc_rng_set_seed(C.byref(r), 34, 2619) == 0
    assert [lib.orc_i4_uni(C.byref(r)) for _ in range(6)] == [2042271690, 2025478916, 143468825, 1820612537,
                                                            694284300, 1065740867]
    lib.orc_rng_set_seed(C.byref(r), 34, 2619)
    got = [f32_bits(lib.orc_r4_uni_01(C.byref(r))) for _ in range(6)]
    assert got == [0x3f737531, 0x3f7174b7, 0x3d88d292, 0x3f5908b0, 0x3ea587c0, 0x3efe17aa]
    # set_seed range checks (rnglib.cpp:1714-1728)
    assert lib.orc_rng_set_seed(C.byref(r), 0, 5) != 0
    assert lib.orc_rng_set_seed(C.byref(r), 5, 2147483399) != 0


@pytest.mark.parametrize("n,p,want,state", [
    (0, 0.03, 0, (1360476, 106572348)),
    (1000, 0.028, 37, (1360476, 106572348)),
    (10000, 0.999, 9985, (1360476, 106572348)),
    (10000000, 4.6e-5, 417, (750997589, 873002235)),
])
def test_ignbin_known_answers(n, p, want, state):
    lib = po.orc()
    r = po.OrcRng()
    lib.orc_rng_set_seed(C.byref(r), 34, 2619)
    assert lib.orc_ignbin(C.byref(r), n, p) == want
    assert (r.cg1, r.cg2) == state


@pytest.mark.parametrize("name", sorted(cases.CASES))
def test_oracle_matches_golden(name, workdir):
    g = util.golden()[name]
    case = cases.CASES[name]()
    assert util.sha(case["fasta"].tobytes()) == g["fasta_sha256"], "synthetic generator drifted"
    r = po.oracle_simulate(util.write_fasta(case, workdir, name), tu=case["tu"], tb=case["tb"], tc=case["tc"],
                           td=case["td"], **cases.sim_kwargs(case))
    assert r.rc == 0
    for i, s in enumerate(r.streams):
        assert len(s) == g["streams"][i]["bytes"], (name, i)
        assert util.sha(s) == g["streams"][i]["sha256"], (name, i)
    assert list(r.cg) == g["cg_exit"]


def test_verbatim_golden_files(workdir):
    case = cases.CASES["inversion"]()
    r = po.oracle_simulate(util.write_fasta(case, workdir, "inversion"), tu=case["tu"], tb=case["tb"], tc=case["tc"],
                           td=case["td"], **cases.sim_kwargs(case))
    for s, nm in zip(r.streams, po.OUT_NAMES):
        with open(os.path.join(util.ROOT, "tests", "golden", f"inversion_{nm}.fastq"), "rb") as fh:
            assert fh.read() == s, nm


@pytest.mark.skipif(not po.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("seed", [1, 17, 4242])
def test_oracle_matches_reference_live(seed, workdir):
    fa = synth.make_fasta([("r1", 200000 + seed), ("r2", 3000)], seed=seed, lead_n=100)
    p = os.path.join(workdir, f"live{seed}.fa")
    fa.tofile(p)
    tu = synth.templates_unpaired(50, 60, seed, max_x=4) + synth.templates_unpaired(20, 90, seed + 1, indel=True)
    tb = synth.templates_badend(20, 60, seed + 2, opp_len=30)
    tc = synth.templates_paired(30, 60, seed + 3, frag_mean=200, frag_sd=20)
    td = synth.templates_paired(30, 60, seed + 4, discordant=True, max_fraglen=3000)
    kw = dict(tu=tu, tb=tb, tc=tc, td=td, fraction=400.0, function=3, min_u=1, min_c=1, min_d=1, min_b=1,
              cg=cases.cli_seed(seed))
    a = po.ref_simulate(p, **kw)
    b = po.oracle_simulate(p, **kw)
    assert a.rc == 0 and b.rc == 0
    assert a.streams == b.streams
    assert a.cg == b.cg


@pytest.mark.skipif(not po.have_ref(), reason="oracle/_ref not built")
def test_ignbin_matches_reference_sweep():
    """orc_ignbin vs the reference's ignbin (ranlib.cpp:1094) over both branches, p on both
    sides of 0.5, value and generator state."""
    o, r = po.orc(), po.ref()
    rng = np.random.default_rng(3)
    for _ in range(400):
        n = int(rng.choice([0, 1, 7, 100, 3000, 30000, 1000000, 10000000]))
        p = float(np.float32(rng.choice([1e-6, 4.6e-5, 1e-3, 0.03, 0.3, 0.5, 0.7, 0.97, 0.999])))
        s = (int(rng.integers(1, 2147483562)), int(rng.integers(1, 2147483398)))
        st = po.OrcRng()
        o.orc_rng_set_seed(C.byref(st), *s)
        r.ref_set_seed(*s)
        for _k in range(5):
            assert o.orc_ignbin(C.byref(st), n, p) == r.ref_ignbin(n, p), (n, p, s)
        a, b = C.c_int(), C.c_int()
        r.ref_get_state(C.byref(a), C.byref(b))
        assert (st.cg1, st.cg2) == (a.value, b.value)


def _one_read(workdir, ref: bytes, qual: bytes, xs: bytes, fw: bool):
    """One unpaired read drawn from a reference just long enough (forces offset 0)."""
    p = os.path.join(workdir, "prop.fa")
    with open(p, "wb") as fh:
        fh.write(b">r1\n" + ref + b"\n")
    t = TemplateUnpaired(0, len(qual), "T" if fw else "F", "0", 0, qual, xs)
    # one window, p clamps to .999 -> essentially always >= 1 read; take the first
    r = po.oracle_simulate(p, tu=[t], fraction=50.0, function=3, min_u=1, min_c=1, min_d=1, min_b=1, cg=(34, 2619),
                           chunksz=4096)
    assert r.rc == 0 and r.wrote[0] >= 1
    lines = r.streams[0].split(b"\n")
    return lines[0], lines[1], lines[3]


def test_simplesim_properties(workdir):
    """simplesim.cpp:414-512 (tests 1-7).  rflen == len(ref)-1 would leave one slot; the
    reference tests call init() directly at offset 0, here the template spans all but the
    last base so nslots == 1 and the drawn offset is always 0."""
    # t1/t2: '=' copies, strand does not change the stored read; fw prints as is
    name, seq, qual = _one_read(workdir, b"ACGTA", b"ABCD", b"====", True)
    assert seq == b"ACGT" and qual == b"ABCD" and name.startswith(b"@qtip!:r1:+:0:0:u")
    # t3: X differs from the reference base
    _, seq, _ = _one_read(workdir, b"ACGTA", b"ABCD", b"=X==", True)
    assert seq[0:1] == b"A" and seq[1:2] != b"C" and seq[2:] == b"GT"
    # t4: D skips a reference base
    _, seq, qual = _one_read(workdir, b"ACGTA", b"ABC", b"=D==", True)
    assert seq == b"AGT" and qual == b"ABC"
    # t5: I inserts a random base
    _, seq, _ = _one_read(workdir, b"AGTA", b"ABCD", b"=I==", True)
    assert seq[0:1] == b"A" and seq[2:] == b"GT"
    # t7: reverse strand prints the reverse complement and reversed qualities
    name, seq, qual = _one_read(workdir, b"AAACCA", b"EDCBA", b"=====", False)
    assert seq == b"GGTTT" and qual == b"ABCDE" and b":-:" in name


def test_fasta_windows_reference_unit_test(workdir):
    """fasta.cpp:134-266, first file of the reference's own test: chunksz 2, olap 1."""
    p = os.path.join(workdir, "ft.fa")
    with open(p, "wb") as fh:
        fh.write(b">record1 ok\nAAAACCCCGGGG\nTTTT\n>record2 mk\nA\nT\n>record3\tblah\nA")
    lib = po.orc()
    fa = lib.orc_fasta_load(p.encode())
    assert lib.orc_fasta_nrecords(fa) == 3
    assert [lib.orc_fasta_name(fa, i) for i in range(3)] == [b"record1", b"record2", b"record3"]
    assert [lib.orc_fasta_name_full(fa, i) for i in range(3)] == [b"record1 ok", b"record2 mk", b"record3\tblah"]
    n = C.c_size_t()
    ch = lib.orc_fasta_chunks(fa, 2, 1, C.byref(n))
    got = [(ch[i].record, ch[i].refoff, ch[i].retsz) for i in range(n.value)]
    want = [(0, i, 2) for i in range(15)] + [(1, 0, 2), (2, 0, 1)]
    assert got == want
    seq0 = C.string_at(lib.orc_fasta_seq(fa, 0, None), 16)
    assert seq0 == b"AAAACCCCGGGGTTTT"
    lib.orc_fasta_free(fa)


def test_apply_function():
    """simplesim.cpp:177-190"""
    lib = po.orc()
    assert lib.orc_apply_function(45.0, 2, 30000, 10000) == 30000  # 45*sqrt(1e4)=4500 < min
    assert lib.orc_apply_function(45.0, 2, 100, 1000000) == 45000
    assert lib.orc_apply_function(0.5, 1, 1, 1000) == 500
    assert lib.orc_apply_function(1234.0, 3, 1, 5) == 1234
    assert lib.orc_apply_function(1234.0, 3, 1, 0) == 0


@pytest.mark.skipif(not po.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("seed", cases.RANDOM_SEEDS)
def test_oracle_matches_reference_random_configs(seed, workdir):
    """Randomised pin (cases.random_case): record layout, window size, target function (sqrt / linear / const with
    `added` counts), read lengths, indel / soft-clipped transcripts and the category mix all drawn from the seed; the
    oracle must produce the unmodified reference's seven streams and exit state byte for byte."""
    case = cases.random_case(seed)
    p = os.path.join(workdir, f"rand{seed}.fa")
    case["fasta"].tofile(p)
    kw = dict(tu=case["tu"], tb=case["tb"], tc=case["tc"], td=case["td"], **cases.sim_kwargs(case))
    a = po.ref_simulate(p, **kw)
    b = po.oracle_simulate(p, **kw)
    assert a.rc == 0 and b.rc == 0
    assert a.streams == b.streams
    assert a.cg == b.cg


# ---- tandem-SAM correctness labelling: the oracle's restatement against the reference's own code ----

def label_cases(seed=5, n=4000):
    """(qname, rname, flag, pos) records: names in the simulator's wire format (simplesim.cpp:102-171) for every read
    type, aligned where they came from or somewhere else, plus wgsim-style and foreign names.  None trips one of the
    reference's asserts (an unpaired FLAG on a pair's name, a name that stops after 'qtip!')."""
    import random
    rnd = random.Random(seed)
    refs = [b"chr1", b"chr10", b"chrX", b"contig_7", b"r", b"chr1_alt"]
    out = []
    for _ in range(n):
        kind = rnd.choice(["u", "u", "c", "d", "b1", "b2", "wgsim", "junk"])
        ref = rnd.choice(refs)
        off = rnd.choice([0, 1, 29, 30, 31, rnd.randint(0, 10**6), rnd.randint(0, 3 * 10**9), 4294967295 + rnd.randint(0, 99)])
        fw = rnd.random() < 0.5
        score = rnd.randint(-300, 60)
        dpos = rnd.choice([0, 0, 0, 1, -1, 29, -29, 30, -30, 31, rnd.randint(-10**5, 10**5)])
        aref = ref if rnd.random() < 0.85 else rnd.choice(refs)
        afw = fw if rnd.random() < 0.85 else not fw
        if kind == "u":
            q = b"qtip!:%s:%s:%d:%d:u" % (ref, b"+" if fw else b"-", off, score)
            if rnd.random() < 0.1:
                q += b" extra"
            flag = 0 if afw else 16
            out.append((q, aref, flag, max(0, off + 1 + dpos)))
        elif kind in ("c", "d", "b1", "b2"):
            off2 = off + rnd.randint(-400, 400) if kind != "b1" and kind != "b2" else off
            off2 = max(0, off2)
            fw2 = (not fw) if kind == "c" else (rnd.random() < 0.5)
            if kind in ("b1", "b2"):
                fw2 = fw
            q = b"qtip!:%s:%s:%d:%d:%s:%s:%d:%d:%s" % (ref, b"+" if fw else b"-", off, score, ref, b"+" if fw2 else b"-", off2,
                                                         rnd.randint(-300, 60), kind.encode())
            mate2 = rnd.random() < 0.5
            want_fw = fw2 if mate2 else fw
            afw = want_fw if rnd.random() < 0.85 else not want_fw
            flag = 1 | (128 if mate2 else 64) | (0 if afw else 16) | (2 if kind == "c" else 0)
            base = off2 if mate2 else off
            out.append((q, aref, flag, max(0, base + 1 + dpos)))
        elif kind == "wgsim":
            fs, l1, l2 = rnd.randint(1, 10**7), rnd.randint(30, 250), rnd.randint(30, 250)
            fe = fs + rnd.randint(250, 800)
            flip = rnd.randint(0, 1)
            q = b"%s_%d_%d_0:0:0_0:0:0_%d_%d_%d_%x" % (ref, fs, fe, l1, l2, flip, rnd.randint(0, 255))
            mate2 = rnd.random() < 0.5
            left = (not flip) == (not mate2)
            truth = fs if left else fe - (l2 if mate2 else l1) + 1
            flag = 1 | (128 if mate2 else 64) | (16 if rnd.random() < 0.5 else 0)
            out.append((q, aref, flag, max(1, truth + dpos)))
        else:
            q = rnd.choice([b"read_17", b"SRR1.55", b"qtip", b"a_b_c_d_e_f_g_h_i", b"x_1_2_3:4:5_6:7:8_9_10_1_a"])
            out.append((q, aref, rnd.choice([0, 16, 65, 129]), rnd.randint(1, 10**6)))
    return out


def test_oracle_labelling_matches_reference_set_correctness():
    """orc_set_correctness == the UNMODIFIED reference's Alignment::set_correctness (oracle/_ref/libreflabel.so) on 4000
    records covering every read type, both mates, strand / reference / position disagreements around the wiggle, wgsim
    names and foreign names, for three wiggle values."""
    if not os.path.exists(po.REF_LABEL_SO):
        pytest.skip("oracle/_ref/libreflabel.so not built (no /root/reference here)")
    o, r = po.orc(), po.ref_label()
    recs = label_cases()
    seen = set()
    for wiggle in (30, 1, 1000):
        for q, rn, flag, pos in recs:
            a = o.orc_set_correctness(q, rn, flag, pos, wiggle)
            b = r.ref_set_correctness(q, rn, flag, pos, wiggle)
            assert a == b, (q, rn, flag, pos, wiggle, a, b)
            seen.add(a)
    assert seen == {-1, 0, 1}


# ---- SAM fields -> edit transcript: the oracle's restatement against the reference's own parser ----

def sam_fields_from_xscript(xs: bytes, rnd):
    """An M-style CIGAR + MD:Z and an extended (=/X) CIGAR that encode the transcript xs (ops = X I D S)."""
    import itertools
    runs = [(k, len(list(g))) for k, g in itertools.groupby(xs.decode())]
    eqx = "".join(f"{n}{op}" for op, n in runs)
    m_runs = []
    for op, n in runs:
        mop = "M" if op in "=X" else op
        if m_runs and m_runs[-1][0] == mop:
            m_runs[-1][1] += n
        else:
            m_runs.append([mop, n])
    cigar = "".join(f"{n}{op}" for op, n in m_runs)
    md, cnt = [], 0
    for op, n in runs:
        if op == "=":
            cnt += n
        elif op == "X":
            for _ in range(n):
                md.append(f"{cnt}{rnd.choice('ACGT')}")
                cnt = 0
        elif op == "D":
            md.append(f"{cnt}^" + "".join(rnd.choice("ACGT") for _ in range(n)))
            cnt = 0
    md.append(str(cnt))
    return cigar.encode(), "".join(md).encode(), eqx.encode()


def xscript_cases(seed=3, n=1500):
    """(cigar, mdz or None, pos, expected transcript) from synthetic templates: substitutions only, indels, soft clips,
    hard clips, both CIGAR styles."""
    import random
    from qtip_b200 import synth
    rnd = random.Random(seed)
    ts = (synth.templates_unpaired(n // 3, 100, seed, max_x=5) + synth.templates_unpaired(n // 3, 150, seed + 1, indel=True, len_jitter=20) +
          synth.templates_unpaired(n // 3, 60, seed + 2, indel=True))
    out = []
    for t in ts:
        xs = t.edit_xscript
        if rnd.random() < 0.3:  # soft clips at the ends
            xs = b"S" * rnd.randint(1, 9) + xs + (b"S" * rnd.randint(1, 9) if rnd.random() < 0.5 else b"")
        cigar, mdz, eqx = sam_fields_from_xscript(xs, rnd)
        pos = rnd.choice([1, 7, rnd.randint(1, 10**6), rnd.randint(1, 3 * 10**9)])
        if rnd.random() < 0.2:  # hard clips leave no trace in the transcript
            cigar = b"%dH" % rnd.randint(1, 30) + cigar + b"%dH" % rnd.randint(1, 30)
        if rnd.random() < 0.5:
            out.append((cigar, mdz, pos, xs))
        else:
            out.append((eqx, mdz if rnd.random() < 0.5 else None, pos, xs))
    return out


def test_oracle_edit_xscript_matches_reference_parser():
    """orc_edit_xscript == the UNMODIFIED reference's parse_cigar / mdz_to_list / cigar_and_mdz_to_edit_xscript / lpos /
    rpos (qtip_parse.cpp:164-186, :223-401; oracle/_ref/libreflabel.so): transcript, clips and both positions, and the
    transcript is the one the CIGAR / MD:Z pair was made from."""
    if not os.path.exists(po.REF_LABEL_SO):
        pytest.skip("oracle/_ref/libreflabel.so not built (no /root/reference here)")
    o, r = po.orc(), po.ref_label()
    n = 0
    for cigar, mdz, pos, xs in xscript_cases():
        a = po.edit_xscript(o.orc_edit_xscript, cigar, mdz, pos)
        b = po.edit_xscript(r.ref_edit_xscript, cigar, mdz, pos)
        assert a == b, (cigar, mdz, pos, a, b)
        assert a[0] == 0 and a[1] == xs, (cigar, mdz, a)
        n += 1
    assert n >= 1400
