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
                                   dict(QTIPSIM_R2_K="1"), dict(QTIPSIM_R2_K="400", QTIPSIM_R2_SIGMA="8")])
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
