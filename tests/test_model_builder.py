"""The input model's reservoirs fed with plain arrays (qtipsim_model_begin / _add / _commit; qtip_b200/csrc/model.cpp),
no GPU: the host logic behind the three entry points against

* the unmodified reference's ReservoirSampledEList::add_part1 (ds.h:437-446, compiled from where it lies into
  oracle/_ref/libref.so): slots returned, draws consumed, generator state afterwards;
* the call sites' rule restated (qtip_parse.cpp:849-861, :1148-1166: `if (off < k) list().back().init(...)`) over four
  reservoirs sharing the one generator in SAM order, driven by the reference's own r4_uni_01;
* qtipsim_set_templates_* fed with the same kept templates: same olap, same replay plan."""
import ctypes as C
import random

import numpy as np
import pytest

import cases
import util
from oracle import pyoracle as po
from qtip_b200 import _abi, synth
from qtip_b200.templates import TemplatePaired, TemplateUnpaired


def hooks():
    L = util.hosttest()
    L.qthost_model_begin.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
    L.qthost_model_add.argtypes = [C.c_void_p, C.POINTER(_abi.ModelBatch), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    L.qthost_model_kept.argtypes = [C.c_void_p, C.c_int]
    L.qthost_model_kept.restype = C.c_uint64
    L.qthost_model_offered.argtypes = [C.c_void_p, C.c_int]
    L.qthost_model_offered.restype = C.c_uint64
    L.qthost_model_record.argtypes = [C.c_void_p, C.c_int, C.c_uint64]
    L.qthost_model_record.restype = C.c_char_p
    L.qthost_model_commit.argtypes = [C.c_void_p]
    return L


def make_records(seed, n, cats=(0, 1, 2, 3)):
    """n records in 'SAM order': (cat, template).  Alignment-level arrays are built by batch_of()."""
    rnd = random.Random(seed)
    tu = synth.templates_unpaired(n, 40, seed, max_x=3)
    tb = synth.templates_badend(n, 40, seed + 1, opp_len=30)
    tc = synth.templates_paired(n, 40, seed + 2, frag_mean=200, frag_sd=20)
    td = synth.templates_paired(n, 40, seed + 3, discordant=True, max_fraglen=2000)
    pools = {0: tu, 1: tb, 2: tc, 3: td}
    return [(c, pools[c][i]) for i, c in enumerate(rnd.choice(cats) for _ in range(n))]


def batch_of(recs):
    """qtipsim_model_batch over recs; returns (batch, keepalive)."""
    score, ln, fw, quals, q_off, xs, x_off = [], [], [], b"", [], b"", []

    def aln(s, l, f, q, x):
        nonlocal quals, xs
        score.append(s), ln.append(l), fw.append(1 if f == "T" else 0)
        q_off.append(len(quals)), x_off.append(len(xs))
        quals += q + b"\0"
        xs += x + b"\0"
        return len(score) - 1

    cat, a1, a2, mf, ol, up, fl = [], [], [], [], [], [], []
    for c, t in recs:
        cat.append(c)
        if c < 2:
            a1.append(aln(t.best_score, t.len, t.fw_flag, t.qual, t.edit_xscript)), a2.append(0)
            mf.append(ord(t.mate_flag)), ol.append(t.opp_len), up.append(0), fl.append(0)
        else:
            a1.append(aln(t.score_1, t.len_1, t.fw_flag_1, t.qual_1, t.edit_xscript_1))
            a2.append(aln(t.score_2, t.len_2, t.fw_flag_2, t.qual_2, t.edit_xscript_2))
            mf.append(ord("0")), ol.append(0), up.append(1 if t.upstream1 else 0), fl.append(t.fraglen)
    arrs = dict(best_score=np.array(score, np.int32), len=np.array(ln, np.int32), fw=np.array(fw, np.uint8),
                q_off=np.array(q_off, np.uint64), x_off=np.array(x_off, np.uint64), cat=np.array(cat, np.uint8),
                a1=np.array(a1, np.uint64), a2=np.array(a2, np.uint64), mate_flag=np.array(mf, np.uint8),
                opp_len=np.array(ol, np.int32), upstream1=np.array(up, np.uint8), fraglen=np.array(fl, np.uint64))
    b = _abi.ModelBatch()
    for k, v in arrs.items():
        setattr(b, k, v.ctypes.data_as(dict(_abi.ModelBatch._fields_)[k]))
    b.quals, b.xscripts, b.n, b.n_aln = quals, xs, len(recs), len(score)
    return b, (arrs, quals, xs)


def record_text(c, t):
    if c < 2:
        return (f"{t.best_score},{t.len},{t.fw_flag},{t.qual.decode()},{t.edit_xscript.decode()}|0,0,T,,|"
                f"{t.mate_flag},{t.opp_len},0,0")
    return (f"{t.score_1},{t.len_1},{t.fw_flag_1},{t.qual_1.decode()},{t.edit_xscript_1.decode()}|"
            f"{t.score_2},{t.len_2},{t.fw_flag_2},{t.qual_2.decode()},{t.edit_xscript_2.decode()}|"
            f"0,0,{1 if t.upstream1 else 0},{t.fraglen}")


def run_builder(L, recs, k, cg, batches=1):
    h = C.c_void_p(L.qthost_create())
    L.qthost_model_begin(h, (C.c_uint64 * 4)(*k))
    cg1, cg2 = C.c_int32(cg[0]), C.c_int32(cg[1])
    step = max(1, (len(recs) + batches - 1) // batches)
    for i in range(0, len(recs), step):
        b, keep = batch_of(recs[i:i + step])
        assert L.qthost_model_add(h, C.byref(b), C.byref(cg1), C.byref(cg2)) == 0
    kept = [[L.qthost_model_record(h, c, i).decode() for i in range(L.qthost_model_kept(h, c))] for c in range(4)]
    offered = [L.qthost_model_offered(h, c) for c in range(4)]
    return h, kept, offered, (cg1.value, cg2.value)


@pytest.mark.skipif(not po.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("k,n,seed", [(1, 50, 3), (5, 5, 4), (5, 6, 5), (30, 2000, 6), (200, 150, 7), (64, 5000, 8), (0, 40, 9)])
def test_reservoir_matches_reference_add_part1(k, n, seed):
    """One reservoir: the slots the reference's add_part1 returns decide what is kept; same draws, same exit state."""
    ref = po.ref()
    ref.ref_reservoir_offsets.argtypes = [C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64)]
    ref.ref_reservoir_offsets.restype = C.c_uint64
    cg = cases.cli_seed(seed)
    ref.ref_set_seed(*cg)
    off = (C.c_uint64 * n)()
    size = ref.ref_reservoir_offsets(k, n, off)
    a, b = C.c_int(), C.c_int()
    ref.ref_get_state(C.byref(a), C.byref(b))
    recs = make_records(seed, n, cats=(0,))
    want = []
    for i, (c, t) in enumerate(recs):  # the call site: if (off < k) list().back().init(...)
        if len(want) < k:
            assert off[i] == len(want)
            want.append(None)
        if off[i] < k:
            want[-1] = record_text(c, t)
    L = hooks()
    h, kept, offered, cg_out = run_builder(L, recs, (k, 0, 0, 0), cg, batches=3)
    assert len(kept[0]) == size == min(k, n)
    assert kept[0] == want
    assert offered == [n, 0, 0, 0]
    assert cg_out == (a.value, b.value)
    L.qthost_destroy(h)


@pytest.mark.skipif(not po.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("seed", range(6))
def test_four_reservoirs_share_the_generator_in_sam_order(seed):
    """Records of all four categories interleaved: the restated call sites driven by the reference's r4_uni_01."""
    ref = po.ref()
    rnd = random.Random(100 + seed)
    k = [rnd.choice([1, 7, 50]) for _ in range(4)]
    recs = make_records(seed, 1500)
    cg = cases.cli_seed(20 + seed)
    ref.ref_set_seed(*cg)
    want, n = [[] for _ in range(4)], [0] * 4
    for c, t in recs:
        n[c] += 1
        if len(want[c]) < k[c]:
            want[c].append(None)
            off = len(want[c]) - 1
        else:
            r4 = np.float32(ref.ref_r4_uni_01())
            off = min(int(r4 * np.float32(n[c])), n[c] - 1)  # (size_t)(r4_uni_01() * n_): a float product
        if off < k[c]:
            want[c][-1] = record_text(c, t)
    a, b = C.c_int(), C.c_int()
    ref.ref_get_state(C.byref(a), C.byref(b))
    L = hooks()
    h, kept, offered, cg_out = run_builder(L, recs, k, cg, batches=7)
    assert kept == want
    assert offered == n
    assert cg_out == (a.value, b.value)
    L.qthost_destroy(h)


def test_commit_equals_set_templates(workdir):
    """The committed model is the model qtipsim_set_templates_* builds from the same kept templates and counts."""
    L = hooks()
    recs = make_records(11, 900)
    k = (40, 25, 60, 10)
    cg = cases.cli_seed(5)
    h, kept, offered, _ = run_builder(L, recs, k, cg, batches=2)
    assert L.qthost_model_commit(h) == 0
    by_text = {record_text(c, t): t for c, t in recs}
    model = [[by_text[s] for s in kept[c]] for c in range(4)]
    fa = synth.make_fasta([("a", 60000), ("b", 20000)], seed=12)
    case = dict(fasta=fa, tu=model[0], tb=model[1], tc=model[2], td=model[3], added=tuple(offered), chunksz=8192,
                kw=dict(fraction=3.0, function=1, min_u=5, min_c=5, min_d=5, min_b=5, cg=cases.cli_seed(6)))
    hs = util.HostSide(case)
    kw = cases.sim_kwargs(case)
    want = hs.plan(kw)
    want_recs = [hs.plan_recs(c, want[3][c]) for c in range(4)]
    # the same genome on the handle the model was committed on
    f = np.ascontiguousarray(fa)
    L.qthost_parse(h, C.c_void_p(f.ctypes.data), f.size)
    hs2 = util.HostSide.__new__(util.HostSide)
    hs2.L, hs2.h, hs2._keep = L, h, []
    assert hs2.L.qthost_olap(hs2.h) == hs.L.qthost_olap(hs.h)
    assert all(L.qthost_cat_valid(h, c) for c in range(4))
    got = hs2.plan(kw)
    assert got == want and want[0] == 0 and sum(want[3]) > 0
    assert [hs2.plan_recs(c, got[3][c]) for c in range(4)] == want_recs
    hs.close()
    hs2.close()


def test_argument_errors():
    L = hooks()
    h = C.c_void_p(L.qthost_create())
    b, keep = batch_of(make_records(1, 10))
    cg1, cg2 = C.c_int32(1), C.c_int32(2)
    assert L.qthost_model_add(h, C.byref(b), C.byref(cg1), C.byref(cg2)) == -7  # before begin: QTIPSIM_ESTATE
    L.qthost_model_begin(h, (C.c_uint64 * 4)(5, 5, 5, 5))
    keep[0]["cat"][3] = 9
    assert L.qthost_model_add(h, C.byref(b), C.byref(cg1), C.byref(cg2)) == -1  # QTIPSIM_EINVAL
    keep[0]["cat"][3] = 0
    keep[0]["a1"][5] = b.n_aln
    assert L.qthost_model_add(h, C.byref(b), C.byref(cg1), C.byref(cg2)) == -1  # alignment index out of range
    L.qthost_destroy(h)


@pytest.mark.gpu
def test_model_through_the_abi_gives_the_same_reads(workdir):
    """qtipsim_model_begin / _add / _commit on a device handle, then qtipsim_run: the seven streams and the exit state of a
    run whose models were set with qtipsim_set_templates_* from the same kept templates (and of the oracle)."""
    from qtip_b200 import sim
    L = hooks()
    recs = make_records(11, 900)
    k = (40, 25, 60, 10)
    cg = cases.cli_seed(5)
    h, kept, offered, cg_out = run_builder(L, recs, k, cg, batches=2)
    L.qthost_destroy(h)
    by_text = {record_text(c, t): t for c, t in recs}
    model = [[by_text[s] for s in kept[c]] for c in range(4)]
    fa = synth.make_fasta([("a", 60000), ("b", 20000)], seed=12)
    case = dict(fasta=fa, tu=model[0], tb=model[1], tc=model[2], td=model[3], added=tuple(offered), chunksz=8192,
                kw=dict(fraction=3.0, function=1, min_u=5, min_c=5, min_d=5, min_b=5, cg=cases.cli_seed(6)))
    want, st_w, p_w = util.run_cuda(case, workdir, "modelabi")
    eng = sim.Engine()
    try:
        eng.load_fasta(util.write_fasta(case, workdir, "modelabi"))
        eng.model_begin(k)
        state = cg
        for i in range(0, len(recs), 450):
            b, keep = batch_of(recs[i:i + 450])
            state = eng.model_add(b, state)
        assert state == cg_out
        got_kept, got_added = eng.model_commit()
        assert got_kept == [len(x) for x in kept] and got_added == offered
        kw = case["kw"]
        p = sim.make_params(case["chunksz"], kw["fraction"], kw["function"], kw["min_u"], kw["min_c"], kw["min_d"], kw["min_b"],
                            0, cg=kw["cg"])
        st = eng.run(p)
        assert eng.results(st) == want
        assert (p.cg1, p.cg2) == (p_w.cg1, p_w.cg2)
        assert list(st.wrote) == list(st_w.wrote) and sum(st.wrote) > 0
    finally:
        eng.close()
    o = po.oracle_simulate(util.write_fasta(case, workdir, "modelabi"), tu=case["tu"], tb=case["tb"], tc=case["tc"], td=case["td"],
                           **cases.sim_kwargs(case))
    assert o.rc == 0 and want == o.streams


def test_sam_fields_to_input_model_pipeline():
    """Both halves of the template construction chained on the host: CIGAR / MD:Z -> edit transcripts (xscript.h, what
    k_xs_lengths / k_xs_fill run per alignment) written into one arena the way qtipsim_edit_xscripts lays it out
    (NUL-terminated, x_off[i], x_off[n] = bytes used), that arena handed to the reservoirs as it is, commit: the model is
    valid, holds the transcripts the SAM fields were made from, and its olap is the longest kept reference span."""
    from test_oracle_pin import xscript_cases
    L = hooks()
    rnd = random.Random(3)
    sam = xscript_cases(seed=21, n=600)
    arena, x_off, score, ln, fw, quals, q_off = b"", [], [], [], [], b"", []
    for cigar, mdz, pos, xs in sam:
        out = C.create_string_buffer(1 << 12)
        lc, rc_, lp, rp = C.c_int(), C.c_int(), C.c_uint64(), C.c_uint64()
        assert L.qthost_edit_xscript(cigar, mdz, pos, out, 1 << 12, C.byref(lc), C.byref(rc_), C.byref(lp), C.byref(rp)) == 0
        x_off.append(len(arena))
        arena += out.value + b"\0"
        rl = sum(out.value.count(c) for c in (b"=", b"X", b"I", b"S"))
        score.append(-rnd.randint(0, 60)), ln.append(rl), fw.append(rnd.randint(0, 1))
        q_off.append(len(quals))
        quals += bytes(rnd.randint(45, 73) for _ in range(rl)) + b"\0"  # (no ',' : the record text below is split on it)
    x_off.append(len(arena))
    n = len(sam)
    arrs = dict(best_score=np.array(score, np.int32), len=np.array(ln, np.int32), fw=np.array(fw, np.uint8),
                q_off=np.array(q_off, np.uint64), x_off=np.array(x_off, np.uint64), cat=np.zeros(n, np.uint8),
                a1=np.arange(n, dtype=np.uint64), mate_flag=np.full(n, ord("0"), np.uint8), opp_len=np.zeros(n, np.int32))
    b = _abi.ModelBatch()
    for k_, v in arrs.items():
        setattr(b, k_, v.ctypes.data_as(dict(_abi.ModelBatch._fields_)[k_]))
    b.quals, b.xscripts, b.n, b.n_aln = quals, arena, n, n
    h = C.c_void_p(L.qthost_create())
    L.qthost_model_begin(h, (C.c_uint64 * 4)(100, 0, 0, 0))
    cg1, cg2 = C.c_int32(11), C.c_int32(22)
    assert L.qthost_model_add(h, C.byref(b), C.byref(cg1), C.byref(cg2)) == 0
    assert (cg1.value, cg2.value) != (11, 22)  # 500 records were offered to the full reservoir
    kept = [L.qthost_model_record(h, 0, i).decode().split("|")[0].split(",") for i in range(L.qthost_model_kept(h, 0))]
    assert len(kept) == 100 and L.qthost_model_offered(h, 0) == n
    made_from = {xs.decode() for _, _, _, xs in sam}
    assert all(k_[4] in made_from for k_ in kept)
    assert [k_[4] for k_ in kept[:99]] == [sam[i][3].decode() for i in range(99)]  # (only the last slot is ever replaced)
    assert L.qthost_model_commit(h) == 0
    assert L.qthost_cat_valid(h, 0) == 1
    L.qthost_olap.restype = C.c_uint64
    assert L.qthost_olap(h) == max(sum(k_[4].count(c) for c in "=XDS") for k_ in kept)
    L.qthost_destroy(h)
