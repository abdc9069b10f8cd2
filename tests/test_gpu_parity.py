"""Parity tests proper: the CUDA path, called through the C ABI (include/qtipsim.h), against the
CPU oracle and the golden fixtures generated from the unmodified reference.

Replay mode must be BIT-EXACT (all seven FASTQ streams, counts, and the generator state handed
back).  Native (Philox) mode is checked through size-independent properties and chi-square
tests on position, strand, template/fragment length and edit spectrum.
"""
import os

import numpy as np
import pytest

import cases
import util

pytestmark = pytest.mark.gpu


def _oracle(case, workdir, name):
    from oracle import pyoracle as po
    return po.oracle_simulate(util.write_fasta(case, workdir, name), tu=case["tu"], tb=case["tb"], tc=case["tc"],
                              td=case["td"], **cases.sim_kwargs(case))


@pytest.mark.parametrize("name", sorted(cases.CASES))
def test_replay_bit_exact_vs_oracle_and_golden(name, workdir):
    case = cases.CASES[name]()
    streams, st, p = util.run_cuda(case, workdir, name)
    o = _oracle(case, workdir, name)
    g = util.golden()[name]
    assert list(st.wrote) == o.wrote
    assert list(st.target) == o.target
    for i in range(7):
        assert len(streams[i]) == len(o.streams[i]), (name, i)
        if streams[i] != o.streams[i]:
            a, b = np.frombuffer(streams[i], np.uint8), np.frombuffer(o.streams[i], np.uint8)
            k = int(np.nonzero(a != b)[0][0])
            lo = max(0, k - 200)
            raise AssertionError(f"{name} stream {i} differs at byte {k}:\nGPU: {streams[i][lo:k + 60]!r}\nORC: {o.streams[i][lo:k + 60]!r}")
        assert util.sha(streams[i]) == g["streams"][i]["sha256"], (name, i)
    assert (p.cg1, p.cg2) == tuple(o.cg) == tuple(g["cg_exit"])
    assert st.draws == o.draws
    assert st.kernel_launches > 0


def test_replay_from_host_bytes_equals_from_path(workdir):
    case = cases.CASES["mixed_small"]()
    a, _, _ = util.run_cuda(case, workdir, "mixed_small", via="path")
    b, _, _ = util.run_cuda(case, workdir, "mixed_small", via="bytes")
    assert a == b


def test_replay_sharded_concatenation_is_identical(workdir):
    """shards emit contiguous slices; rank-order concatenation == single run."""
    case = cases.CASES["c4_mix_small"]()
    full, st, pf = util.run_cuda(case, workdir, "c4_mix_small")
    for nshard in (2, 3):
        parts, cg = [], None
        for r in range(nshard):  # the chain is sequential: rank r starts from rank r-1's exit state
            s, _, p = util.run_cuda(case, workdir, "c4_mix_small", shard=(r, nshard), cg=cg)
            cg = (p.cg1, p.cg2)
            parts.append(s)
        for i in range(7):
            assert b"".join(p[i] for p in parts) == full[i], (nshard, i)
        assert cg == (pf.cg1, pf.cg2)


def test_python_mirror_of_reference_interface(workdir):
    """StreamingSimulator/InputModel* mirror (qtip_b200/sim.py): same call sequence as
    qtip_parse.cpp:1825-1867, files written, global generator advanced."""
    from qtip_b200 import rnglib, sim
    case = cases.CASES["mixed_small"]()
    o = _oracle(case, workdir, "mixed_small")
    fa = util.write_fasta(case, workdir, "mixed_small")
    mu = sim.InputModelUnpaired(case["tu"], len(case["tu"]))
    mb = sim.InputModelUnpaired(case["tb"], len(case["tb"]))
    mc = sim.InputModelPaired(case["tc"], len(case["tc"]))
    md = sim.InputModelPaired(case["td"], len(case["td"]))
    paths = [os.path.join(workdir, f"mirror_reads_{n}.fastq") for n in sim._abi.OUT_NAMES]
    fhs = [open(p, "wb") for p in paths]
    rnglib.initialize()
    rnglib.set_seed(*case["kw"]["cg"])
    ss = sim.StreamingSimulator([fa], 128 * 1024, mu, mb, mc, md, *fhs)
    assert ss.num_estimated_bases() == case["fasta"].size
    kw = case["kw"]
    ss.simulate_batch(kw["fraction"], kw["function"], kw["min_u"], kw["min_c"], kw["min_d"], kw["min_b"])
    for fh in fhs:
        fh.close()
    for p, want in zip(paths, o.streams):
        with open(p, "rb") as fh:
            assert fh.read() == want
    assert rnglib.get_state() == tuple(o.cg)


def test_error_behaviour(workdir):
    """Where the reference throws / asserts, the C ABI returns an error code."""
    from qtip_b200 import sim
    from qtip_b200.templates import TemplateUnpaired
    eng = sim.Engine()
    with pytest.raises(sim.QtipSimError) as ei:
        eng.load_fasta(os.path.join(workdir, "does_not_exist.fa"))
    assert ei.value.code == -4  # fasta.cpp:54-58
    case = cases.CASES["inversion"]()
    eng.load_fasta(util.write_fasta(case, workdir, "inversion"))
    with pytest.raises(sim.QtipSimError) as ei:  # nothing to simulate: simplesim.cpp:209
        eng.run(sim.make_params(cg=(1, 1)))
    assert ei.value.code == -6
    eng.set_unpaired(0, [TemplateUnpaired(0, 4, "T", "0", 0, b"IIII", b"==N=")], 1)
    with pytest.raises(sim.QtipSimError) as ei:  # unknown op: simplesim.cpp:58-60
        eng.run(sim.make_params(cg=(1, 1)))
    assert ei.value.code == -5
    eng.set_unpaired(0, case["tu"], len(case["tu"]))
    with pytest.raises(sim.QtipSimError) as ei:  # chunksz <= olap: fasta.h:41
        eng.run(sim.make_params(chunksz=60, cg=(1, 1)))
    assert ei.value.code == -6
    with pytest.raises(sim.QtipSimError) as ei:  # set_seed range: rnglib.cpp:1714-1728
        eng.run(sim.make_params(cg=(0, 1)))
    assert ei.value.code == -6
    eng.close()


# ------------------------------------------------------------------ native (Philox) mode --

def _parse_fastq(b: bytes):
    lines = b.split(b"\n")
    assert lines[-1] == b""
    n = (len(lines) - 1) // 4
    return [(lines[4 * i], lines[4 * i + 1], lines[4 * i + 3]) for i in range(n)]


def _revcomp(s: bytes) -> bytes:
    return s.translate(bytes.maketrans(b"ACGTN", b"TGCAN"))[::-1]


def _check_records_against_genome(case, recs, genome, typ):
    """Every read is its template applied to the reference at the position its name claims:
    '=' columns equal the reference, X columns differ from it (simplesim.cpp:38-62)."""
    tmpl_by_key = {}
    for t in case["tu"]:
        key = (t.fw_flag == "T", t.best_score, t.qual if t.fw_flag == "T" else t.qual[::-1])
        tmpl_by_key.setdefault(key, []).append(t)
    nchecked = 0
    for name, seq, qual in recs:
        f = name.split(b":")
        assert f[0] == b"@qtip!" and f[-1] == typ
        refid, strand, off, score = f[1], f[2], int(f[3]), int(f[4])
        fw = strand == b"+"
        cands = tmpl_by_key.get((fw, score, qual))
        assert cands, name
        read = seq if fw else _revcomp(seq)
        ref = genome[refid]
        ok = False
        for t in cands:
            rd = rf = 0
            good = True
            for op in t.edit_xscript:
                c = chr(op)
                if c == "=":
                    good &= read[rd] == ref[off + rf]
                    rd += 1; rf += 1
                elif c == "X":
                    good &= read[rd] != ref[off + rf] and read[rd] in b"ACGT"
                    rd += 1; rf += 1
                elif c == "I":
                    rd += 1
                elif c == "D":
                    rf += 1
                elif c == "S":
                    rd += 1; rf += 1
                if not good:
                    break
            if good and rd == len(read):
                ok = True
                break
        assert ok, name
        nchecked += 1
    return nchecked


def _genome_dict(case, workdir, name):
    import ctypes as C
    from oracle import pyoracle as po
    lib = po.orc()
    fa = lib.orc_fasta_load(util.write_fasta(case, workdir, name).encode())
    out = {}
    for i in range(lib.orc_fasta_nrecords(fa)):
        L = C.c_uint64()
        p = lib.orc_fasta_seq(fa, i, C.byref(L))
        out[lib.orc_fasta_name(fa, i)] = C.string_at(p, L.value)
    lib.orc_fasta_free(fa)
    return out


@pytest.mark.parametrize("name", ["c3_shape_small", "c5_indel", "edges"])
def test_native_reads_are_consistent_with_genome(name, workdir):
    case = cases.CASES[name]()
    streams, st, _ = util.run_cuda(case, workdir, name, rng_mode=1, philox_seed=12345)
    recs = _parse_fastq(streams[0])
    assert len(recs) == st.wrote[0] > 0
    n = _check_records_against_genome(case, recs, _genome_dict(case, workdir, name), b"u")
    assert n == len(recs)


def _xscript_matches(xs: bytes, read: bytes, ref: bytes, off: int) -> bool:
    """read = transcript xs applied to ref at off (simplesim.cpp:38-62): '=' copies, X differs (any base over an N),
    I / S are free draws, D skips a reference base.  ref is padded with N past the record end."""
    rd = rf = 0
    for op in xs:
        c = chr(op)
        if c == "=":
            if read[rd] != ref[off + rf]:
                return False
            rd += 1; rf += 1
        elif c == "X":
            if read[rd] == ref[off + rf] or read[rd] not in b"ACGT":
                return False
            rd += 1; rf += 1
        elif c == "I":
            if read[rd] not in b"ACGT":
                return False
            rd += 1
        elif c == "D":
            rf += 1
        elif c == "S":
            if read[rd] not in b"ACGT":
                return False
            rd += 1; rf += 1
    return rd == len(read)


def _reflen(xs: bytes) -> int:
    return sum(xs.count(c) for c in (b"S", b"=", b"X", b"D"))


def _check_pairs_against_genome(case, cat, recs1, recs2, genome):
    """Native-mode pairs against the genome at the offsets their (shared) name claims.
    c / d (simplesim.cpp:342-397): mate 1 in file _1 and mate 2 in file _2 are some template's two transcripts applied at
    off_1 / off_2, and the two offsets obey SURVEY Q9 (simplesim.cpp:354-363): upstream1 => off_2 - off_1 =
    max(fraglen, rflen_2) - rflen_2, else the mirror image.  b (simplesim.cpp:279-334, simplesim.h:89-117): the aligned
    mate (file _1 for b1, _2 for b2) is a template applied at the named offset; the other mate is opp_len random ACGT with
    quality 'I' throughout; both halves of the name repeat the aligned mate's origin."""
    assert len(recs1) == len(recs2) > 0
    pad = {k: v + b"N" * 70000 for k, v in genome.items()}
    orient = lambda q, fw: q if fw else q[::-1]  # noqa: E731
    if cat == "b":
        by = {}
        for t in case["tb"]:
            by.setdefault((t.fw_flag == "T", t.best_score, orient(t.qual, t.fw_flag == "T"), t.mate_flag), []).append(t)
    else:
        by = {}
        for t in case["tc" if cat == "c" else "td"]:
            f1, f2 = t.fw_flag_1 == "T", t.fw_flag_2 == "T"
            by.setdefault((f1, t.score_1, orient(t.qual_1, f1), f2, t.score_2, orient(t.qual_2, f2)), []).append(t)
    for (n1, s1, q1), (n2, s2, q2) in zip(recs1, recs2):
        assert n1 == n2, (n1, n2)  # write_pair prints the same name on both mates (simplesim.cpp:132-171)
        f = n1.split(b":")
        assert f[0] == b"@qtip!" and len(f) == 10, n1
        ref1, fw1, off1, sc1 = f[1], f[2] == b"+", int(f[3]), int(f[4])
        ref2, fw2, off2, sc2 = f[5], f[6] == b"+", int(f[7]), int(f[8])
        typ = f[9]
        if cat == "b":
            assert typ in (b"b1", b"b2") and (ref1, fw1, off1, sc1) == (ref2, fw2, off2, sc2), n1
            aligned_is_1 = typ == b"b1"
            sa, qa = (s1, q1) if aligned_is_1 else (s2, q2)
            sr, qr = (s2, q2) if aligned_is_1 else (s1, q1)
            cands = by.get((fw1, sc1, qa, "1" if aligned_is_1 else "2"))
            assert cands, n1
            read = sa if fw1 else _revcomp(sa)
            assert any(_xscript_matches(t.edit_xscript, read, pad[ref1], off1) and len(sr) == t.opp_len for t in cands), n1
            assert qr == b"I" * len(sr) and set(sr) <= set(b"ACGT"), n1
        else:
            assert typ == cat.encode() and ref1 == ref2, n1
            cands = by.get((fw1, sc1, q1, fw2, sc2, q2))
            assert cands, n1
            r1 = s1 if fw1 else _revcomp(s1)
            r2 = s2 if fw2 else _revcomp(s2)
            ok = False
            for t in cands:
                rl1, rl2 = _reflen(t.edit_xscript_1), _reflen(t.edit_xscript_2)
                if t.upstream1:
                    if off2 - off1 != max(t.fraglen, rl2) - rl2:
                        continue
                elif off1 - off2 != max(t.fraglen, rl1) - rl1:
                    continue
                if _xscript_matches(t.edit_xscript_1, r1, pad[ref1], off1) and _xscript_matches(t.edit_xscript_2, r2, pad[ref2], off2):
                    ok = True
                    break
            assert ok, n1
    return len(recs1)


@pytest.mark.parametrize("name", ["c4_mix_small", "mixed_small", "short_reads", "long_refid"])
def test_native_pairs_are_consistent_with_genome(name, workdir):
    """Every native-mode bad-end / concordant / discordant pair, both files, against the genome and the templates."""
    case = cases.CASES[name]()
    streams, st, _ = util.run_cuda(case, workdir, name, rng_mode=1, philox_seed=424242)
    genome = _genome_dict(case, workdir, name)
    checked = 0
    for cat, i1, ci in (("b", 1, 1), ("c", 3, 2), ("d", 5, 3)):
        if not st.wrote[ci]:
            continue
        r1, r2 = _parse_fastq(streams[i1]), _parse_fastq(streams[i1 + 1])
        assert len(r1) == st.wrote[ci]
        checked += _check_pairs_against_genome(case, cat, r1, r2, genome)
    assert checked > 0
    if st.wrote[0]:
        assert _check_records_against_genome(case, _parse_fastq(streams[0]), genome, b"u") == st.wrote[0]


def test_native_is_deterministic_and_seed_sensitive(workdir):
    case = cases.CASES["mixed_small"]()
    a, _, _ = util.run_cuda(case, workdir, "mixed_small", rng_mode=1, philox_seed=7)
    b, _, _ = util.run_cuda(case, workdir, "mixed_small", rng_mode=1, philox_seed=7)
    c, _, _ = util.run_cuda(case, workdir, "mixed_small", rng_mode=1, philox_seed=8)
    assert a == b
    assert a != c


def test_native_sharding_is_invariant(workdir):
    """Counter-based draws are keyed by the global read index: any shard count gives the same bytes."""
    case = cases.CASES["c4_mix_small"]()
    full, _, _ = util.run_cuda(case, workdir, "c4_mix_small", rng_mode=1, philox_seed=99)
    parts = [util.run_cuda(case, workdir, "c4_mix_small", rng_mode=1, philox_seed=99, shard=(r, 4))[0] for r in range(4)]
    for i in range(7):
        assert b"".join(p[i] for p in parts) == full[i]


def _chi2_ok(obs_a, obs_b, alpha=1e-4):
    """Two-sample chi-square homogeneity test on two histograms over the same bins."""
    from scipy import stats
    a, b = np.asarray(obs_a, float), np.asarray(obs_b, float)
    keep = (a + b) >= 10
    a, b = a[keep], b[keep]
    if a.size < 2:
        return True
    _, p, _, _ = stats.chi2_contingency(np.vstack([a, b]))
    return p > alpha


def test_native_distributionally_equivalent_to_replay(workdir):
    """north_star: native-RNG mode is distributionally equivalent to the reference's stream by
    chi-square on position, strand, template (=> fragment length, quality) and edit spectrum."""
    case = cases.CASES["c3_shape_small"]()
    case["kw"] = dict(case["kw"], fraction=60000.0)
    rep, st_r, _ = util.run_cuda(case, workdir, "c3_shape_small")
    nat, st_n, _ = util.run_cuda(case, workdir, "c3_shape_small", rng_mode=1, philox_seed=2024)
    ra, rb = _parse_fastq(rep[0]), _parse_fastq(nat[0])
    # read count: both ~ sum of Binomial(n, p_w); compare within 5 sigma of each other
    assert abs(len(ra) - len(rb)) < 5 * np.sqrt(len(ra) + len(rb))
    genome = _genome_dict(case, workdir, "c3_shape_small")
    names = sorted(genome)

    def feats(recs):
        rec_idx, pos_bin, strand, xcount, subst = [], [], [], [], np.zeros((4, 4))
        for name, seq, qual in recs:
            f = name.split(b":")
            refid, fw, off = f[1], f[2] == b"+", int(f[3])
            rec_idx.append(names.index(refid))
            pos_bin.append(min(9, off * 10 // max(1, len(genome[refid]))))
            strand.append(int(fw))
            read = seq if fw else _revcomp(seq)
            ref = genome[refid][off:off + len(read)]  # these templates have no indels
            a = np.frombuffer(read, np.uint8)
            b = np.frombuffer(ref, np.uint8)
            mm = np.nonzero(a != b)[0]
            xcount.append(len(mm))
            for k in mm:
                if ref[k:k + 1] in b"ACGT":
                    subst[b"ACGT".index(ref[k:k + 1]), b"ACGT".index(read[k:k + 1])] += 1
        return (np.bincount(rec_idx, minlength=len(names)), np.bincount(pos_bin, minlength=10),
                np.bincount(strand, minlength=2), np.bincount(xcount, minlength=8), subst)

    fa, fb = feats(ra), feats(rb)
    assert _chi2_ok(fa[0], fb[0]), "record (position) distribution"
    assert _chi2_ok(fa[1], fb[1]), "within-record position distribution"
    assert _chi2_ok(fa[2], fb[2]), "strand"
    assert _chi2_ok(fa[3], fb[3]), "mismatches per read (edit spectrum)"
    assert _chi2_ok(fa[4].reshape(-1), fb[4].reshape(-1)), "substitution matrix"
    # substituted base is uniform over the three alternatives in both modes
    for f in (fa, fb):
        for r in range(4):
            alt = np.delete(f[4][r], r)
            from scipy import stats
            assert stats.chisquare(alt).pvalue > 1e-4


def test_native_pair_mix_distributionally_equivalent_to_replay(workdir):
    """north_star's chi-square criterion on the pair categories (BASELINE config 4 shape, 70/15/15 c/d/b): category counts,
    fragment length, strand combination, position, mismatches per mate, which mate of a bad-end pair is the aligned one and
    the base composition of its random mate -- native (Philox) against replay (the reference's own stream)."""
    from scipy import stats
    case = cases.CASES["c4_mix_small"]()
    case["kw"] = dict(case["kw"], fraction=5.0)  # linear function: 5 x added => ~50 k pairs
    rep, st_r, _ = util.run_cuda(case, workdir, "c4_mix_small")
    nat, st_n, _ = util.run_cuda(case, workdir, "c4_mix_small", rng_mode=1, philox_seed=777)
    for c in range(4):  # counts ~ sum over windows of Binomial(target, p_w) in both modes
        assert abs(st_r.wrote[c] - st_n.wrote[c]) <= 5 * np.sqrt(st_r.wrote[c] + st_n.wrote[c] + 1)
    genome = _genome_dict(case, workdir, "c4_mix_small")
    names = sorted(genome)

    def pair_feats(streams, i1, cat):
        r1, r2 = _parse_fastq(streams[i1]), _parse_fastq(streams[i1 + 1])
        frag, strands, rec, pos, mm, typ, comp = [], [], [], [], [], [], np.zeros(4)
        for (n1, s1, _), (_, s2, _) in zip(r1, r2):
            f = n1.split(b":")
            fw1, off1, fw2, off2 = f[2] == b"+", int(f[3]), f[6] == b"+", int(f[7])
            ref = genome[f[1]]
            rec.append(names.index(f[1]))
            pos.append(min(9, off1 * 10 // max(1, len(ref))))
            strands.append(2 * int(fw1) + int(fw2))
            if cat == "b":
                aligned_is_1 = f[9] == b"b1"
                typ.append(int(aligned_is_1))
                sa, sr = (s1, s2) if aligned_is_1 else (s2, s1)
                a = np.frombuffer(sa if fw1 else _revcomp(sa), np.uint8)
                b = np.frombuffer((ref + b"N" * 1000)[off1:off1 + len(a)], np.uint8)
                mm.append(int((a != b).sum()))
                comp += np.array([sr.count(x) for x in (b"A", b"C", b"G", b"T")])
            else:
                frag.append(min(79, (max(off1, off2) + 150 - min(off1, off2)) // (10 if cat == "c" else 300)))
                for sq, fw, off in ((s1, fw1, off1), (s2, fw2, off2)):  # these templates carry X only: mismatches = X columns
                    a = np.frombuffer(sq if fw else _revcomp(sq), np.uint8)
                    b = np.frombuffer((ref + b"N" * 70000)[off:off + len(a)], np.uint8)
                    mm.append(int((a != b).sum()))
        return dict(frag=np.bincount(frag, minlength=80) if frag else np.zeros(2), strands=np.bincount(strands, minlength=4),
                    rec=np.bincount(rec, minlength=len(names)), pos=np.bincount(pos, minlength=10),
                    mm=np.bincount(np.minimum(mm, 9), minlength=10), typ=np.bincount(typ, minlength=2) if typ else np.zeros(2), comp=comp)

    for cat, i1 in (("b", 1), ("c", 3), ("d", 5)):
        fr, fn = pair_feats(rep, i1, cat), pair_feats(nat, i1, cat)
        for k in ("frag", "strands", "rec", "pos", "mm", "typ"):
            assert _chi2_ok(fr[k], fn[k]), (cat, k)
        if cat == "b":  # init_random draws uniform bases in both modes (simplesim.h:108-113)
            for f in (fr, fn):
                assert stats.chisquare(f["comp"]).pvalue > 1e-4


@pytest.mark.parametrize("name", ["mixed_small", "c5_indel", "long_refid"])
def test_generic_kernel_path_still_bit_exact(name, workdir, monkeypatch):
    """The warp-per-read generic kernel (fallback for records too large to stage) stays covered."""
    monkeypatch.setenv("QTIPSIM_KERNEL", "generic")
    case = cases.CASES[name]()
    streams, st, p = util.run_cuda(case, workdir, name)
    o = _oracle(case, workdir, name)
    assert streams == o.streams
    nat_g, _, _ = util.run_cuda(case, workdir, name, rng_mode=1, philox_seed=5)
    monkeypatch.delenv("QTIPSIM_KERNEL")
    nat_t, _, _ = util.run_cuda(case, workdir, name, rng_mode=1, philox_seed=5)
    assert nat_g == nat_t  # both kernels consume the Philox stream identically


@pytest.mark.parametrize("name", ["mixed_small", "c5_indel", "c4_mix_small", "edges"])
def test_staged_kernel_path_still_bit_exact(name, workdir, monkeypatch):
    """The staged warp-tile kernel (taken by categories with bodies under 17 bytes) stays covered on the
    ordinary cases too."""
    monkeypatch.setenv("QTIPSIM_KERNEL", "staged")
    case = cases.CASES[name]()
    streams, st, p = util.run_cuda(case, workdir, name)
    o = _oracle(case, workdir, name)
    assert streams == o.streams
    nat_s, _, _ = util.run_cuda(case, workdir, name, rng_mode=1, philox_seed=5)
    monkeypatch.delenv("QTIPSIM_KERNEL")
    nat_d, _, _ = util.run_cuda(case, workdir, name, rng_mode=1, philox_seed=5)
    assert nat_s == nat_d  # both kernels consume the Philox stream identically


@pytest.mark.parametrize("name", ["mixed_small", "c5_indel", "c4_mix_small", "edges", "long_refid", "c3_shape_small", "refid_lengths"])
def test_direct_kernel_path_still_bit_exact(name, workdir, monkeypatch):
    """The first-generation direct kernel (simulate_direct.cu: taken when a tile of 16 records does not fit the tile
    kernel's shared-memory budget) stays covered on the ordinary cases, and consumes the Philox stream like the tile
    kernel that replaced it as the default."""
    monkeypatch.setenv("QTIPSIM_KERNEL", "direct")
    case = cases.CASES[name]()
    streams, st, p = util.run_cuda(case, workdir, name)
    o = _oracle(case, workdir, name)
    assert streams == o.streams
    nat_d, _, _ = util.run_cuda(case, workdir, name, rng_mode=1, philox_seed=5)
    monkeypatch.delenv("QTIPSIM_KERNEL")
    nat_t, _, _ = util.run_cuda(case, workdir, name, rng_mode=1, philox_seed=5)
    assert nat_d == nat_t


@pytest.mark.parametrize("name", ["mixed_small", "c4_mix_small", "c3_shape_small", "edges", "long_refid", "short_reads"])
def test_native_single_pass_plan_equals_split_plan(name, workdir, monkeypatch):
    """Native mode computes plan, record lengths and stream offsets in one kernel (single-pass prefix sum with
    decoupled look-back, k_plan_philox_scan); QTIPSIM_PLAN_SPLIT=1 takes the plan + lengths kernel and the three-kernel
    scans it replaced.  Same bytes, same counts."""
    case = cases.CASES[name]()
    a, sta, _ = util.run_cuda(case, workdir, name, rng_mode=1, philox_seed=11)
    monkeypatch.setenv("QTIPSIM_PLAN_SPLIT", "1")
    b, stb, _ = util.run_cuda(case, workdir, name, rng_mode=1, philox_seed=11)
    assert a == b and list(sta.wrote) == list(stb.wrote)


@pytest.mark.parametrize("resolver", ["cluster", "cta"])
@pytest.mark.parametrize("name", ["mixed_small", "c4_mix_small", "inversion", "edges", "c3_shape_small"])
def test_first_generation_resolvers_equal_band_resolver(name, resolver, workdir, monkeypatch):
    """K2 has three forms: the band-speculative staged resolver (default, resolve2.cu), the thread-block-cluster
    resolver and the single-CTA one it grew out of (QTIPSIM_RESOLVER=cluster|cta).  Same plan, same draws, same
    exit state."""
    case = cases.CASES[name]()
    a, sta, pa = util.run_cuda(case, workdir, name)
    monkeypatch.setenv("QTIPSIM_RESOLVER", resolver)
    b, stb, pb = util.run_cuda(case, workdir, name)
    assert a == b and sta.draws == stb.draws and (pa.cg1, pa.cg2) == (pb.cg1, pb.cg2)


@pytest.mark.parametrize("knobs", [dict(QTIPSIM_R2_RCAP="1024", QTIPSIM_R2_RCAP_EMIT="1024"),      # many rounds per window
                                   dict(QTIPSIM_R2_SIGMA="0.3", QTIPSIM_R2_K="16"),                  # bands that miss
                                   dict(QTIPSIM_R2_NBMAX="64", QTIPSIM_R2_K="7"),                    # clipped bands
                                   dict(QTIPSIM_R2_K="1"), dict(QTIPSIM_R2_K="400", QTIPSIM_R2_SIGMA="8"),
                                   dict(QTIPSIM_R2_WALK_MAX="0"), dict(QTIPSIM_R2_WALK_MAX="40")])                   # which emit kernel takes a region
@pytest.mark.parametrize("name", ["mixed_small", "c4_mix_small", "c3_shape_small", "edges", "long_reads", "c5_indel"])
def test_band_resolver_is_exact_under_any_tuning(name, knobs, workdir, monkeypatch):
    """The band resolver's tables are exact or INVALID whatever the band width, the dense-table size or the stage
    size: tiny tables force the round-by-round exact mode, narrow bands force early stage ends, and the result is the
    oracle's every time."""
    case = cases.CASES[name]()
    o = _oracle(case, workdir, name)
    monkeypatch.setenv("QTIPSIM_R2_FORCE", "1")  # also where the resolver would decline (bands that cannot fit a table)
    for k, v in knobs.items():
        monkeypatch.setenv(k, v)
    got, st, p = util.run_cuda(case, workdir, name)
    for i in range(7):
        assert got[i] == o.streams[i], (name, knobs, i)
    assert (p.cg1, p.cg2) == tuple(o.cg) and st.draws == o.draws and list(st.wrote) == list(o.wrote)


def test_host_pack_path_equals_device_pack(workdir, monkeypatch):
    """K0 on the device (default) and the host FASTA packer produce the same reads."""
    case = cases.CASES["edges"]()
    dev, _, _ = util.run_cuda(case, workdir, "edges")
    monkeypatch.setenv("QTIPSIM_HOST_PACK", "1")
    host, _, _ = util.run_cuda(case, workdir, "edges")
    assert dev == host


@pytest.mark.parametrize("name", ["mixed_small", "c4_mix_small", "edges"])
def test_host_planner_path_still_bit_exact(name, workdir, monkeypatch):
    """The host chain walk (QTIPSIM_PLANNER=host) and the device resolver (default) agree."""
    case = cases.CASES[name]()
    dev, st_d, p_d = util.run_cuda(case, workdir, name)
    monkeypatch.setenv("QTIPSIM_PLANNER", "host")
    host, st_h, p_h = util.run_cuda(case, workdir, name)
    assert dev == host
    assert (p_d.cg1, p_d.cg2) == (p_h.cg1, p_h.cg2)
    assert st_d.draws == st_h.draws


@pytest.mark.gpu
@pytest.mark.parametrize("chunk,min_part", [(4096, 1), (65536, 100000), (0, 0)])
@pytest.mark.parametrize("name", ["mixed_small", "edges", "c3_shape_small", "c4_mix_small"])
def test_streaming_path_equals_one_shot_path(name, chunk, min_part, workdir, monkeypatch):
    """qtipsim_stream_begin / qtipsim_stream_run (upload, pack, simulate and download overlapped, record by record) give
    the bytes of load + run: replay mode against the oracle / golden, native mode against the one-shot path; with
    upload chunks small enough that records span chunks and every case runs in several parts."""
    if chunk:
        monkeypatch.setenv("QTIPSIM_STREAM_CHUNK", str(chunk))
        monkeypatch.setenv("QTIPSIM_STREAM_MIN_PART", str(min_part))
    case = cases.CASES[name]()
    want, st0, p0 = util.run_cuda(case, workdir, name, via="bytes")
    got, st, p, again = util.run_cuda_streaming(case)
    for i in range(7):
        assert got[i] == want[i], (name, i)
        assert again[i] == want[i], (name, i)
    assert (p.cg1, p.cg2) == (p0.cg1, p0.cg2)
    assert list(st.wrote) == list(st0.wrote) and list(st.bytes) == list(st0.bytes)
    wantn, _, _ = util.run_cuda(case, workdir, name, rng_mode=1, philox_seed=77, via="bytes")
    gotn, _, _, againn = util.run_cuda_streaming(case, rng_mode=1, philox_seed=77)
    for i in range(7):
        assert gotn[i] == wantn[i], (name, i)
        assert againn[i] == wantn[i], (name, i)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", cases.RANDOM_SEEDS)
def test_replay_bit_exact_random_configs(seed, workdir):
    """The randomised configurations the oracle is pinned on (test_oracle_pin.py) through the CUDA path, replay mode:
    all seven streams and the generator state equal the oracle's."""
    from oracle import pyoracle as po
    case = cases.random_case(seed)
    name = f"rand{seed}"
    o = po.oracle_simulate(util.write_fasta(case, workdir, name), tu=case["tu"], tb=case["tb"], tc=case["tc"],
                           td=case["td"], **cases.sim_kwargs(case))
    assert o.rc == 0
    got, st, p = util.run_cuda(case, workdir, name, via="bytes")
    for i in range(7):
        assert got[i] == o.streams[i], (seed, i)
    assert (p.cg1, p.cg2) == tuple(o.cg)
    assert list(st.wrote) == list(o.wrote)


# ------------------------------------------------------------------ tandem-SAM correctness labelling --

def test_labelling_kernel_on_our_own_read_names(workdir):
    """The read names the simulate kernels emit, fed back -- as the aligner's SAM would -- to k_label_correctness
    through the C ABI: an alignment at the named origin is labelled correct for every read type and both mates, shifted
    / flipped / moved ones are labelled like the reference's Alignment::set_correctness (qtip_parse.cpp:519-713) labels
    them (oracle restatement always, the unmodified reference where oracle/_ref/libreflabel.so is present)."""
    import random
    from oracle import pyoracle as po
    from qtip_b200 import sim
    case = cases.CASES["mixed_small"]()
    streams, st, _ = util.run_cuda(case, workdir, "mixed_small")
    rnd = random.Random(3)
    q, r, fl, ps, truth = [], [], [], [], []
    for si, mate in ((0, 0), (1, 1), (2, 2), (3, 1), (4, 2), (5, 1), (6, 2)):
        for name, _, _ in _parse_fastq(streams[si])[:600]:
            f = name[1:].split(b":")
            h = 0 if mate != 2 else 4  # which half of the name describes this mate
            ref, fw, off = f[1 + h], f[2 + h] == b"+", int(f[3 + h])
            base = (0 if mate == 0 else (1 | (64 if mate == 1 else 128)))
            for variant in ("true", "shift", "flip", "other"):
                pos, afw, aref = off + 1, fw, ref
                if variant == "shift":
                    pos = max(1, off + 1 + rnd.choice([-31, -30, -29, 29, 30, 31, 5000]))
                elif variant == "flip":
                    afw = not fw
                elif variant == "other":
                    aref = b"chrB" if ref != b"chrB" else b"chrA"
                q.append(name[1:]); r.append(aref); fl.append(base | (0 if afw else 16)); ps.append(pos); truth.append(variant == "true")
    eng = sim.Engine()
    try:
        got = eng.label_correctness(q, r, fl, ps, wiggle=30)
    finally:
        eng.close()
    o = po.orc()
    want = np.array([o.orc_set_correctness(a, b, c, d, 30) for a, b, c, d in zip(q, r, fl, ps)], dtype=np.int8)
    assert (got == want).all()
    assert all(got[i] == 1 for i in range(len(q)) if truth[i])  # the wire format round-trips
    assert (got == 0).sum() > len(q) // 3
    if os.path.exists(po.REF_LABEL_SO):
        rl = po.ref_label()
        ref = np.array([rl.ref_set_correctness(a, b, c, d, 30) for a, b, c, d in zip(q, r, fl, ps)], dtype=np.int8)
        assert (got == ref).all()


def test_labelling_kernel_matches_oracle_on_generated_records():
    from oracle import pyoracle as po
    from qtip_b200 import sim
    from test_oracle_pin import label_cases
    recs = label_cases(seed=21, n=20000)
    eng = sim.Engine()
    try:
        for wiggle in (30, 2):
            got = eng.label_correctness([x[0] for x in recs], [x[1] for x in recs], [x[2] for x in recs], [x[3] for x in recs], wiggle)
            o = po.orc()
            want = np.array([o.orc_set_correctness(a, b, c, d, wiggle) for a, b, c, d in recs], dtype=np.int8)
            assert (got == want).all()
            assert set(np.unique(got)) == {-1, 0, 1}
    finally:
        eng.close()


# ------------------------------------------------------------------ SAM fields -> edit transcripts --

def test_edit_xscript_kernels_match_oracle():
    """k_xs_lengths / k_xs_fill through qtipsim_edit_xscripts against the oracle's restatement of the reference's CIGAR /
    MD:Z parser on 6000 alignments (both CIGAR styles, indels, soft and hard clips, 64-bit positions) and on malformed
    records; the transcripts of well-formed records are the ones their SAM fields were generated from."""
    from oracle import pyoracle as po
    from qtip_b200 import sim
    from test_oracle_pin import xscript_cases
    recs = xscript_cases(seed=44, n=6000)
    bad = [(b"10M", None, 5, None), (b"5M2D5M", b"10", 9, None), (b"3P10M", b"10", 1, None), (b"M10", b"10", 1, None),
           (b"4=1X5M", None, 3, None), (b"10M", b"5AC3", 1, None)]
    allr = recs + bad
    eng = sim.Engine()
    try:
        got = eng.edit_xscripts([r[0] for r in allr], [r[1] for r in allr], [r[2] for r in allr])
    finally:
        eng.close()
    o = po.orc()
    for (cigar, mdz, pos, xs), g in zip(allr, got):
        want = po.edit_xscript(o.orc_edit_xscript, cigar, mdz, pos)
        assert g[0] == want[0], (cigar, mdz, g, want)
        if want[0] == 0:
            assert g == want, (cigar, mdz, g, want)
            assert xs is None or g[1] == xs
    assert sum(1 for g in got if g[0] != 0) >= 5  # the malformed records (one of the six above is in fact well-formed)


# ------------------------------------------------------------------ packed-genome cache --

def test_packed_genome_cache_round_trip(workdir):
    """qtipsim_load_fasta_cached: the first load packs the FASTA and writes the cache, the second uploads the cache
    (no parse, no pack kernels) and simulates the same bytes; a FASTA that changed on disk invalidates the cache."""
    import time
    from qtip_b200 import sim

    def run(case, fa_path, cache):
        eng = sim.Engine()
        try:
            hit = eng.load_fasta_cached(fa_path, cache)
            added = case.get("added") or (len(case["tu"]), len(case["tb"]), len(case["tc"]), len(case["td"]))
            eng.set_unpaired(0, case["tu"], added[0]); eng.set_unpaired(1, case["tb"], added[1])
            eng.set_paired(2, case["tc"], added[2]); eng.set_paired(3, case["td"], added[3])
            kw = case["kw"]
            p = sim.make_params(case.get("chunksz", 128 * 1024), kw["fraction"], kw["function"], kw["min_u"], kw["min_c"], kw["min_d"],
                                kw["min_b"], 0, cg=kw["cg"])
            st = eng.run(p)
            return hit, eng.results(st), st.kernel_launches
        finally:
            eng.close()

    case = cases.CASES["mixed_small"]()
    fa_path = os.path.join(workdir, "cache_case.fa")
    cache = os.path.join(workdir, "cache_case.qtpg")
    for p_ in (fa_path, cache):
        if os.path.exists(p_):
            os.unlink(p_)
    case["fasta"].tofile(fa_path)
    o = _oracle(case, workdir, "mixed_small")
    hit1, s1, l1 = run(case, fa_path, cache)
    assert not hit1 and os.path.getsize(cache) > 0
    hit2, s2, l2 = run(case, fa_path, cache)
    assert hit2 and s1 == s2 == o.streams
    # another genome under the same name: the stamp (size, mtime) no longer matches
    case2 = cases.CASES["edges"]()
    time.sleep(0.01)
    case2["fasta"].tofile(fa_path)
    o2 = _oracle(case2, workdir, "edges")
    hit3, s3, _ = run(case2, fa_path, cache)
    assert not hit3 and s3 == o2.streams
    hit4, s4, _ = run(case2, fa_path, cache)
    assert hit4 and s4 == o2.streams


# ------------------------------------------------------------------ output paths to the aligner --

@pytest.mark.timeout(180)
def test_interleaved_and_fifo_outputs(workdir):
    """qtipsim_result_interleaved: mate 1 / mate 2 records of every pair category alternate in one stream, byte for byte the
    interleave of the two files.  qtipsim_result_stream_fds: all seven streams into pipes whose reader consumes each
    category's _1 and _2 in lock step, record by record (what an aligner fed through FIFOs does) -- the sequential
    writer would deadlock there -- and the bytes are the files'."""
    import threading
    from qtip_b200 import sim
    case = cases.CASES["c4_mix_small"]()
    case["kw"] = dict(case["kw"], fraction=3.0)
    o = _oracle(case, workdir, "c4_mix_small")
    eng = sim.Engine()
    try:
        eng.load_fasta_bytes(case["fasta"])
        added = case.get("added")
        eng.set_unpaired(0, case["tu"], added[0]); eng.set_unpaired(1, case["tb"], added[1])
        eng.set_paired(2, case["tc"], added[2]); eng.set_paired(3, case["td"], added[3])
        kw = case["kw"]
        st = eng.run(sim.make_params(128 * 1024, kw["fraction"], kw["function"], kw["min_u"], kw["min_c"], kw["min_d"], kw["min_b"], 0, cg=kw["cg"]))
        streams = eng.results(st)
        assert streams == o.streams
        for cat in (1, 2, 3):
            a, b = _split_records(streams[2 * cat - 1]), _split_records(streams[2 * cat])
            want = b"".join(x + y for x, y in zip(a, b))
            assert eng.result_interleaved(cat, st) == want and len(want) > 100000
        # FIFO-style consumers
        pipes = [os.pipe() for _ in range(7)]
        got = [bytearray() for _ in range(7)]

        def lockstep(i1, i2):
            f1, f2 = os.fdopen(pipes[i1][0], "rb"), os.fdopen(pipes[i2][0], "rb")
            while True:
                r1 = [f1.readline() for _ in range(4)]
                r2 = [f2.readline() for _ in range(4)]
                if not r1[0] and not r2[0]:
                    break
                got[i1] += b"".join(r1); got[i2] += b"".join(r2)
            f1.close(); f2.close()

        def single(i):
            with os.fdopen(pipes[i][0], "rb") as f:
                got[i] += f.read()

        readers = [threading.Thread(target=single, args=(0,))] + [threading.Thread(target=lockstep, args=(i, i + 1)) for i in (1, 3, 5)]
        for t in readers:
            t.start()
        eng.stream_fds([p[1] for p in pipes])
        for p in pipes:
            os.close(p[1])
        for t in readers:
            t.join(60)
            assert not t.is_alive()
        assert [bytes(g) for g in got] == streams
    finally:
        eng.close()


def _split_records(b: bytes):
    lines = b.split(b"\n")
    return [b"\n".join(lines[4 * i:4 * i + 4]) + b"\n" for i in range((len(lines) - 1) // 4)]
