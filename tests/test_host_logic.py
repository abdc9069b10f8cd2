"""Host logic of libqtipsim (no GPU): FASTA parse + 2-bit pack, window table, template
flattening, the addressable L'Ecuyer stream, ignbin set-up/sampler, glibc-exact logf and
the replay chain walk -- each against the oracle."""
import ctypes as C

import numpy as np
import pytest

import cases
import util
from oracle import pyoracle as po


@pytest.mark.parametrize("name", ["edges", "mixed_small", "c3_shape_small"])
def test_parse_and_pack_match_oracle(name, workdir):
    case = cases.CASES[name]()
    hs = util.HostSide(case)
    lib = po.orc()
    fa = lib.orc_fasta_load(util.write_fasta(case, workdir, name).encode())
    recs = hs.records()
    assert len(recs) == lib.orc_fasta_nrecords(fa)
    for i, (nm, ln, gs) in enumerate(recs):
        assert nm == lib.orc_fasta_name(fa, i)
        L = C.c_uint64()
        p = lib.orc_fasta_seq(fa, i, C.byref(L))
        assert ln == L.value
        assert gs % 64 == 0
        assert hs.unpack(gs, ln) == C.string_at(p, ln)
        # padding after every record is N
        assert hs.unpack(gs + ln, 1) == b"N"
    # windows
    olap = int(hs.L.qthost_olap(hs.h))
    chunksz = case.get("chunksz", 128 * 1024)
    n = C.c_size_t()
    ch = lib.orc_fasta_chunks(fa, chunksz, olap, C.byref(n))
    want = [(ch[i].record, ch[i].refoff, ch[i].retsz, ch[i].n_count) for i in range(n.value)]
    assert hs.windows(chunksz, olap) == want
    lib.orc_fasta_free(fa)
    hs.close()


def test_stream_is_addressable():
    """state_k = state_0 * a^k: jumping equals stepping (rnglib.cpp:645-659)."""
    L = util.hosttest()
    lib = po.orc()
    r = po.OrcRng()
    lib.orc_rng_set_seed(C.byref(r), 34, 2619)
    s1, s2 = C.c_uint32(34), C.c_uint32(2619)
    i4, r4 = C.c_int32(), C.c_float()
    for k in range(1, 3000):
        L.qthost_rng_step(C.byref(s1), C.byref(s2), C.byref(i4), C.byref(r4))
        want = lib.orc_i4_uni(C.byref(r))
        assert i4.value == want
        if k in (1, 2, 31, 32, 33, 1000, 2999):
            j1, j2 = C.c_uint32(34), C.c_uint32(2619)
            L.qthost_rng_jump(C.byref(j1), C.byref(j2), k)
            assert (j1.value, j2.value) == (s1.value, s2.value) == (r.cg1, r.cg2)
    j1, j2 = C.c_uint32(34), C.c_uint32(2619)
    L.qthost_rng_jump(C.byref(j1), C.byref(j2), 10**12 + 7)
    k1, k2 = C.c_uint32(34), C.c_uint32(2619)
    L.qthost_rng_jump(C.byref(k1), C.byref(k2), 10**12)
    L.qthost_rng_jump(C.byref(k1), C.byref(k2), 7)
    assert (j1.value, j2.value) == (k1.value, k2.value)


def test_ignbin_split_matches_oracle():
    """binom_setup + binom_sample (lecuyer.h) == orc_ignbin, value and state."""
    L = util.hosttest()
    o = po.orc()
    rng = np.random.default_rng(11)
    for _ in range(600):
        n = int(rng.choice([0, 1, 7, 100, 3000, 30000, 1000000, 10000000, 100000000]))
        p = float(np.float32(rng.choice([1e-6, 4.6e-5, 1e-3, 0.03, 0.3, 0.5, 0.7, 0.97, 0.999])))
        s = (int(rng.integers(1, 2147483562)), int(rng.integers(1, 2147483398)))
        st = po.OrcRng()
        o.orc_rng_set_seed(C.byref(st), *s)
        s1, s2 = C.c_uint32(s[0]), C.c_uint32(s[1])
        for _k in range(4):
            assert L.qthost_ignbin(C.byref(s1), C.byref(s2), n, p) == o.orc_ignbin(C.byref(st), n, p), (n, p, s)
        assert (s1.value, s2.value) == (st.cg1, st.cg2)


def test_ignbin_table_sampler_matches_oracle():
    """binom_tables + binom_sample_tab (the device resolver's sampler: BTPE's step-5 bounds read from per-set-up
    tables) == orc_ignbin, value and state, with many variates per set-up so that every branch is taken."""
    L = util.hosttest()
    o = po.orc()
    rng = np.random.default_rng(12)
    for _ in range(120):
        n = int(rng.choice([100, 3000, 30000, 1000000, 10000000, 100000000, 2000000000]))
        p = float(np.float32(rng.choice([4.6e-5, 1e-3, 0.03, 0.3, 0.5, 0.7, 0.97, 0.999])))
        s = (int(rng.integers(1, 2147483562)), int(rng.integers(1, 2147483398)))
        st = po.OrcRng()
        o.orc_rng_set_seed(C.byref(st), *s)
        s1, s2 = C.c_uint32(s[0]), C.c_uint32(s[1])
        for _k in range(300):
            assert L.qthost_ignbin_tab(C.byref(s1), C.byref(s2), n, p) == o.orc_ignbin(C.byref(st), n, p), (n, p, s)
        assert (s1.value, s2.value) == (st.cg1, st.cg2)


def test_draw_base_integer_thresholds():
    """lbase_code_z (three integer comparisons on i4_uni's output) == draw_base through r4_uni_01 (simplesim.h:26-28,
    rnglib.cpp:1502): the thresholds are re-derived by bisection and the two forms compared on +-200 000 around each
    and on a stride of 977 over the whole range."""
    L = util.hosttest()
    t = (C.c_int32 * 3)()
    assert L.qthost_base_thresholds(t, 200000, 977) == 0
    assert list(t) == [536870865, 1073741729, 1610612672]


def test_logf_matches_libm_sampled():
    """logf_glibc == host logf on a stride over all positive normal floats, and densely over
    (0, 1] where ignbin's arguments live.  (Every float: tests/test_logf_exhaustive.py.)"""
    L = util.hosttest()
    assert L.qthost_logf_sweep(0x00800000, 0x7f800000, 4099) == 0
    assert L.qthost_logf_sweep(0x30000000, 0x3f800001, 53) == 0


@pytest.mark.parametrize("name", sorted(cases.CASES))
def test_replay_plan_matches_oracle_trace(name, workdir):
    """The chain walk reproduces, per read, the template and position the reference drew,
    and hands the generator back in the same state after the same number of draws."""
    case = cases.CASES[name]()
    kw = cases.sim_kwargs(case)
    r = po.oracle_simulate(util.write_fasta(case, workdir, name), tu=case["tu"], tb=case["tb"], tc=case["tc"],
                           td=case["td"], want_trace=True, **kw)
    assert r.rc == 0
    hs = util.HostSide(case)
    rc, cg, draws, nreads = hs.plan(kw)
    assert rc == 0
    assert cg == tuple(r.cg)
    assert draws == r.draws
    assert nreads == r.wrote
    wins = hs.windows(kw["chunksz"], int(hs.L.qthost_olap(hs.h)))
    for cat in range(4):
        want = [(t[1], wins[t[2]][0], wins[t[2]][1] + t[3]) for t in r.trace if t[0] == cat]
        got = [(p[0], p[1], p[2]) for p in hs.plan_recs(cat, nreads[cat])]
        assert got == want, (name, cat)
    hs.close()


@pytest.mark.parametrize("seed", cases.RANDOM_SEEDS)
def test_replay_plan_matches_oracle_trace_random_configs(seed, workdir):
    """The product's host logic (FASTA parse, window table, template flattening, replay chain walk) on the randomised
    configurations the oracle is pinned on against the unmodified reference (test_oracle_pin.py)."""
    case = cases.random_case(seed)
    kw = cases.sim_kwargs(case)
    name = f"rand{seed}"
    r = po.oracle_simulate(util.write_fasta(case, workdir, name), tu=case["tu"], tb=case["tb"], tc=case["tc"],
                           td=case["td"], want_trace=True, **kw)
    assert r.rc == 0
    hs = util.HostSide(case)
    rc, cg, draws, nreads = hs.plan(kw)
    assert rc == 0
    assert cg == tuple(r.cg)
    assert draws == r.draws
    assert nreads == r.wrote
    wins = hs.windows(kw["chunksz"], int(hs.L.qthost_olap(hs.h)))
    for cat in range(4):
        want = [(t[1], wins[t[2]][0], wins[t[2]][1] + t[3]) for t in r.trace if t[0] == cat]
        got = [(p[0], p[1], p[2]) for p in hs.plan_recs(cat, nreads[cat])]
        assert got == want, (seed, cat)
    hs.close()


def test_sharded_plans_concatenate(workdir):
    """Shards walk the same chain and emit disjoint, ordered slices of the plan."""
    case = cases.CASES["c3_shape_small"]()
    kw = cases.sim_kwargs(case)
    hs = util.HostSide(case)
    _, _, _, full_n = hs.plan(kw)
    full = hs.plan_recs(0, full_n[0])
    nwin = len(hs.windows(kw["chunksz"], int(hs.L.qthost_olap(hs.h))))
    _, full_cg, full_draws, _ = hs.plan(kw)
    parts, cg, draws = [], None, 0
    for r in range(3):
        lo, hi = r * nwin // 3, (r + 1) * nwin // 3
        _, cg, d, n = hs.plan(kw, lo, hi, cg=cg)  # rank r starts from rank r-1's exit state
        draws += d
        parts += hs.plan_recs(0, n[0])
    assert parts == full
    assert cg == full_cg and draws == full_draws
    hs.close()


def test_bad_transcripts_flagged():
    """Unknown op (simplesim.cpp:58-60) and length mismatch (:63) are detected up front."""
    from qtip_b200.templates import TemplateUnpaired
    bad_op = dict(fasta=np.frombuffer(b">r\nACGT\n", dtype=np.uint8), tu=[TemplateUnpaired(0, 4, "T", "0", 0, b"IIII", b"==N=")],
                  tb=[], tc=[], td=[])
    hs = util.HostSide(bad_op)
    assert hs.L.qthost_cat_valid(hs.h, 0) == 0
    hs.close()
    bad_len = dict(fasta=np.frombuffer(b">r\nACGT\n", dtype=np.uint8), tu=[TemplateUnpaired(0, 4, "T", "0", 0, b"IIII", b"===")],
                   tb=[], tc=[], td=[])
    hs = util.HostSide(bad_len)
    assert hs.L.qthost_cat_valid(hs.h, 0) == 0
    hs.close()


# ---- K0's word-parallel byte classification (qtip_b200/csrc/pack_swar.h) ----

def _byte_class(c: int) -> int:
    """dna_upper + isspace of the reference's parser (fasta.cpp:18-35,113): 0..3 = ACGT, 4 = N, 5 = whitespace."""
    if c in b"Aa":
        return 0
    if c in b"Cc":
        return 1
    if c in b"Gg":
        return 2
    if c in b"Tt":
        return 3
    if c == 0x20 or 9 <= c <= 13:
        return 5
    return 4


def _swar_vec(L, data16: bytes, a: int, b: int):
    codes, nbits, ws = C.c_uint32(), C.c_uint32(), C.c_uint32()
    buf = (C.c_uint8 * 16).from_buffer_copy(data16)
    cnt = L.qthost_swar_vec(buf, a, b, C.byref(codes), C.byref(nbits), C.byref(ws))
    return cnt, codes.value, nbits.value, ws.value


def _want_vec(data16: bytes, a: int, b: int):
    cls = [_byte_class(c) for c in data16]
    ws = sum(1 << k for k in range(16) if cls[k] == 5)
    kept = [c for c in cls[a:b] if c != 5]
    codes = sum((c if c < 4 else 0) << (2 * j) for j, c in enumerate(kept))
    nbits = sum(1 << j for j, c in enumerate(kept) if c == 4)
    return len(kept), codes, nbits, ws


def test_swar_classification_every_byte_value_every_position():
    L = util.hosttest()
    L.qthost_swar_vec.restype = C.c_uint32
    for c in range(256):
        for k in range(16):
            v = bytearray(b"ACGTacgtNn\n\t xR-")  # neighbours of every class around the byte under test
            v[k] = c
            assert _swar_vec(L, bytes(v), 0, 16) == _want_vec(bytes(v), 0, 16), (c, k)


def test_swar_squeeze_random_vectors_and_ranges():
    L = util.hosttest()
    L.qthost_swar_vec.restype = C.c_uint32
    rng = np.random.default_rng(2024)
    alphabet = np.frombuffer(b"ACGTacgtNnRYKM-*\n\r\t \x0b\x0c\x00\x7f\x80\xa0\xc1\xff>", dtype=np.uint8)
    for it in range(4000):
        p_ws = (0.0, 0.05, 0.3, 0.9)[it % 4]
        v = alphabet[rng.integers(0, alphabet.size, 16)].copy()
        v[rng.random(16) < p_ws] = (0x0a, 0x20, 0x0d, 0x09)[it % 4]
        a = int(rng.integers(0, 16))
        b = int(rng.integers(a + 1, 17))
        if it % 3 == 0:
            a, b = 0, 16
        assert _swar_vec(L, v.tobytes(), a, b) == _want_vec(v.tobytes(), a, b), (v.tobytes(), a, b)


def test_swar_pack_unit_arithmetic_matches_bytewise_pack():
    """The bit-offset / image / ship arithmetic of k_pack_units, run lane by lane on the host, for units at every
    text phase and genome phase, against a byte-at-a-time pack."""
    L = util.hosttest()
    L.qthost_swar_pack_unit.restype = C.c_uint64
    L.qthost_swar_pack_unit.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]
    rng = np.random.default_rng(7)
    letters = np.frombuffer(b"ACGTACGTACGTacgtNnRY", dtype=np.uint8)
    for it in range(300):
        n = int(rng.integers(1, 4097)) if it % 5 else 4096
        width = (60, 70, 1, 17, 4096)[it % 5]
        body = letters[rng.integers(0, letters.size, n)]
        lines = [body[i:i + width].tobytes() for i in range(0, n, width)]
        text = (b"\r\n" if it % 7 == 0 else b"\n").join(lines)[:4096]
        if it % 11 == 0:
            text = b" \n\n" + text[:4000] + b"\n\n"
        lead = int(rng.integers(0, 40))
        g = 64 + int(rng.integers(0, 64))
        raw = np.full(lead + len(text) + 600, ord(">"), dtype=np.uint8)  # bytes outside [t0, t1) must be ignored
        raw[lead:lead + len(text)] = np.frombuffer(text, dtype=np.uint8)
        buf = np.zeros(raw.size + 16, dtype=np.uint8)
        off = (-buf.ctypes.data) % 16
        buf[off:off + raw.size] = raw
        base_ptr = buf.ctypes.data + off
        cls = [_byte_class(c) for c in text]
        kept = [c for c in cls if c != 5]
        nb = len(kept)
        bases = np.zeros((g + nb) // 16 + 4, dtype=np.uint32)
        nmask = np.zeros((g + nb) // 32 + 4, dtype=np.uint32)
        # neighbours' bits already in the shared first / last words must survive
        bases[g >> 4] = (1 << (2 * (g & 15))) - 1
        got = L.qthost_swar_pack_unit(base_ptr, lead, lead + len(text), g, bases.ctypes.data, nmask.ctypes.data)
        assert got == nb
        wb = np.zeros_like(bases)
        wn = np.zeros_like(nmask)
        wb[g >> 4] = (1 << (2 * (g & 15))) - 1
        for j, c in enumerate(kept):
            p = g + j
            if c < 4:
                wb[p >> 4] |= np.uint32(c << (2 * (p & 15)))
            else:
                wn[p >> 5] |= np.uint32(1 << (p & 31))
        assert np.array_equal(bases, wb) and np.array_equal(nmask, wn), it


@pytest.mark.parametrize("name", ["edges", "mixed_small"] + [f"rand{s}" for s in cases.RANDOM_SEEDS[:10]] +
                         [f"fuzz{s}" for s in range(24)] + [f"fuzzeof{s}" for s in range(6)])
def test_swar_pack_whole_fasta_matches_host_packer(name):
    """K0's unit arithmetic over whole FASTA files (4 KiB units cut anywhere in a line, CRLF, IUPAC, blank lines,
    headerless leading bases): every record's sequence lines packed unit by unit with the kernel's arithmetic on
    the host equal the host packer's record (which is pinned against the oracle above)."""
    if name.startswith("fuzz"):  # malformed-but-legal text (test_fasta_fuzz.py), incl. files that end inside a header line
        import test_fasta_fuzz
        case = test_fasta_fuzz.fuzz_case(int(name.lstrip("fuzeo")), eof_in_header=name.startswith("fuzzeof"))
    else:
        case = cases.random_case(int(name[4:])) if name.startswith("rand") else cases.CASES[name]()
    hs = util.HostSide(case)
    L = hs.L
    L.qthost_swar_pack_unit.restype = C.c_uint64
    L.qthost_swar_pack_unit.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]
    text = bytes(np.ascontiguousarray(case["fasta"]).tobytes())
    n = len(text)
    # header lines as the loader resolves them: '>' outside a header line starts one, it ends at '\n' or '\r'
    segs, i, hdr_end = [], 0, 0
    first = text.find(b">")
    lead_end = n if first < 0 else first
    if any(_byte_class(c) != 5 for c in text[:lead_end]):
        segs.append([0, lead_end])
    pos = first
    while pos >= 0:
        if pos >= hdr_end:
            if segs:
                segs[-1][1] = min(segs[-1][1], pos)
            j = pos
            while j < n and text[j] not in b"\n\r":
                j += 1
            hdr_end = j
            segs.append([j, n])
        pos = text.find(b">", pos + 1)
    recs = hs.records()
    assert len(segs) == len(recs)
    buf = np.full(n + 64 + 16, 0x20, dtype=np.uint8)
    off = (-buf.ctypes.data) % 16
    buf[off:off + n] = np.frombuffer(text, dtype=np.uint8)
    base_ptr = buf.ctypes.data + off
    total = max(gs + ln for _, ln, gs in recs) + 128
    bases = np.zeros(total // 16 + 8, dtype=np.uint32)
    nmask = np.zeros(total // 32 + 8, dtype=np.uint32)
    for (s0, s1), (nm, ln, gs) in zip(segs, recs):
        g = gs
        for t0 in range(s0, s1, 4096):
            got = L.qthost_swar_pack_unit(base_ptr, t0, min(t0 + 4096, s1), g, bases.ctypes.data, nmask.ctypes.data)
            assert got != 2 ** 64 - 1
            g += got
        assert g - gs == ln, nm
        p = np.arange(gs, gs + ln, dtype=np.int64)
        codes = (bases[p >> 4] >> (2 * (p & 15)).astype(np.uint32)) & 3
        isn = (nmask[p >> 5] >> (p & 31).astype(np.uint32)) & 1
        seq = np.where(isn == 1, ord("N"), np.frombuffer(b"ACGT", dtype=np.uint8)[codes]).astype(np.uint8).tobytes()
        assert seq == hs.unpack(gs, ln), nm
        assert not np.any((codes != 0) & (isn == 1)), nm  # N positions carry code 0, as the kernels expect
    hs.close()


def test_template_compilation_in_slices_equals_single_thread(workdir, monkeypatch):
    """Large models are compiled by several threads in slices that are concatenated (templates.cpp build_category):
    the host plan -- every read's template, position and generator state -- is the one the single-threaded build gives."""
    from qtip_b200 import synth
    fa = synth.make_fasta([("a", 300000)], seed=909)
    case = dict(fasta=fa, tu=synth.templates_unpaired(5000, 100, 910, indel=True, len_jitter=10),
                tb=synth.templates_badend(4500, 80, 911, opp_len=30),
                tc=synth.templates_paired(4200, 100, 912, frag_mean=300, frag_sd=30, indel=True),
                td=synth.templates_paired(4100, 60, 913, discordant=True, max_fraglen=3000))
    kw = dict(fraction=2000.0, function=3, min_u=1, min_c=1, min_d=1, min_b=1, cg=(34, 2619), chunksz=128 * 1024)
    res = []
    for nt in ("1", "5"):
        monkeypatch.setenv("QTIPSIM_TEMPLATE_THREADS", nt)
        hs = util.HostSide(case)
        rc, cg, draws, n = hs.plan(kw)
        assert rc == 0
        res.append((cg, draws, n, [hs.plan_recs(c, min(n[c], 400)) for c in range(4)], int(hs.L.qthost_olap(hs.h))))
        hs.close()
    assert res[0] == res[1]


def test_labelling_function_matches_oracle():
    """label.h::label_correctness (the function k_label_correctness runs per alignment), host build, against the oracle's
    restatement of Alignment::set_correctness (qtip_parse.cpp:519-713) on the shared record generator."""
    from test_oracle_pin import label_cases
    L, o = util.hosttest(), po.orc()
    n1 = 0
    for wiggle in (30, 3):
        for q, rn, flag, pos in label_cases(seed=9, n=3000):
            a = L.qthost_label(q, rn, flag, pos, wiggle)
            assert a == o.orc_set_correctness(q, rn, flag, pos, wiggle), (q, rn, flag, pos, wiggle)
            n1 += a == 1
    assert n1 > 500


def test_edit_xscript_builder_matches_oracle():
    """xscript.h::xs_build / xs_span (what k_xs_lengths / k_xs_fill run per alignment), host build, against the oracle's
    restatement of the reference parser -- well-formed alignments and the malformed ones the reference asserts on."""
    from test_oracle_pin import xscript_cases
    L, o = util.hosttest(), po.orc()
    recs = [(c, m, p) for c, m, p, _ in xscript_cases(seed=8)]
    recs += [(b"10M", None, 5), (b"10M", b"4A6", 5), (b"10M", b"9", 5), (b"5M2D5M", b"5^A5", 9), (b"5M2D5M", b"10", 9), (b"3P10M", b"10", 1),
             (b"M10", b"10", 1), (b"10", b"10", 1), (b"4=1X5M", None, 3), (b"10=", None, 3), (b"5S10=", None, 30), (b"20S", b"0", 4),
             (b"10M5N10M", b"20", 100), (b"10M", b"5AC3", 1), (b"3M1I3M", b"2T3", 77)]
    for cigar, mdz, pos in recs:
        out = C.create_string_buffer(1 << 12)
        lc, rc_ = C.c_int(), C.c_int()
        lp, rp = C.c_uint64(), C.c_uint64()
        st = L.qthost_edit_xscript(cigar, mdz, pos, out, 1 << 12, C.byref(lc), C.byref(rc_), C.byref(lp), C.byref(rp))
        want = po.edit_xscript(o.orc_edit_xscript, cigar, mdz, pos, cap=1 << 12)
        assert st == want[0], (cigar, mdz, st, want)
        if st == 0:
            assert (out.value, lc.value, rc_.value, lp.value, rp.value) == want[1:], (cigar, mdz)
