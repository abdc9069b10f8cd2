"""Randomised odd edit transcripts (no GPU): the oracle against the unmodified reference, and the product's host logic
(template flattening, validity checks, replay chain walk) against the oracle.

Transcripts are drawn from sub-alphabets of "=XIDS" as random strings or as runs of 1-8 equal ops, 1 to 60 ops long:
all-X, all-I (no reference span at all), all-S, leading / trailing / adjacent I and D, single-base reads; unpaired,
bad-end (opp_len 1 / 5 / 40, and now and then 0: assert(len > 0), simplesim.h:96) and concordant templates with fragment
lengths from 1 (shorter than either mate, SURVEY Q9 offsets) to 300.  The reference runs in a forked child: where one of
its asserts aborts the process the oracle must return an error, and the product must refuse the model as well."""
import multiprocessing as mp
import os
import random

import pytest

import cases
import util
from oracle import pyoracle as po
from qtip_b200 import synth
from qtip_b200.templates import TemplatePaired, TemplateUnpaired


def odd_xscript(rnd):
    k = rnd.random()
    n = rnd.choice([1, 2, 3, 5, 17, 30, 60])
    alpha = (b"=X" if k < 0.2 else b"=XID" if k < 0.4 else b"=XIDS" if k < 0.6 else b"X" if k < 0.7 else b"I" if k < 0.8
             else b"D=" if k < 0.9 else b"S")
    if rnd.random() < 0.5:
        out = b""
        while len(out) < n:
            out += bytes([rnd.choice(alpha)]) * rnd.randint(1, 8)
        return out[:n]
    return bytes(rnd.choice(alpha) for _ in range(n))


def read_len(xs):
    return sum(xs.count(c) for c in (b"=", b"X", b"I", b"S"))


def fuzz_model(seed):
    rnd = random.Random(seed)

    def qual(n):
        return bytes(rnd.randint(33, 73) for _ in range(n))

    tu, tb, tc = [], [], []
    for _ in range(30):
        xs = odd_xscript(rnd)
        tu.append(TemplateUnpaired(best_score=-rnd.randint(0, 60), len=read_len(xs), fw_flag=rnd.choice("TF"), mate_flag="0",
                                   opp_len=0, qual=qual(read_len(xs)), edit_xscript=xs))
    for _ in range(20):
        xs = odd_xscript(rnd)
        tb.append(TemplateUnpaired(best_score=-rnd.randint(0, 60), len=read_len(xs), fw_flag=rnd.choice("TF"),
                                   mate_flag=rnd.choice("12"), opp_len=rnd.choice([1, 5, 40]) if rnd.random() < 0.98 else 0,
                                   qual=qual(read_len(xs)), edit_xscript=xs))
    for _ in range(20):
        x1, x2 = odd_xscript(rnd), odd_xscript(rnd)
        tc.append(TemplatePaired(score_12=-rnd.randint(0, 99), score_1=-rnd.randint(0, 60), len_1=read_len(x1),
                                 fw_flag_1=rnd.choice("TF"), qual_1=qual(read_len(x1)), edit_xscript_1=x1,
                                 score_2=-rnd.randint(0, 60), len_2=read_len(x2), fw_flag_2=rnd.choice("TF"),
                                 qual_2=qual(read_len(x2)), edit_xscript_2=x2, upstream1=rnd.random() < 0.5,
                                 fraglen=rnd.choice([1, 10, 60, 61, 100, 300])))
    fa = synth.make_fasta([("a", 30000), ("b", 9000)], seed=5)
    return dict(fasta=fa, tu=tu, tb=tb, tc=tc, td=[], chunksz=rnd.choice([512, 4096]),
                kw=dict(fraction=150.0, function=3, min_u=1, min_c=1, min_d=1, min_b=1, cg=cases.cli_seed(seed + 1)))


def _reference_child(conn, path, kw):
    import faulthandler
    faulthandler.disable()  # an assert in the reference is an expected outcome here
    os.dup2(os.open(os.devnull, os.O_WRONLY), 2)
    r = po.ref_simulate(path, **kw)
    conn.send((r.rc, r.streams, tuple(r.cg)))
    conn.close()


@pytest.mark.parametrize("seed", range(24))
def test_odd_transcripts(seed, workdir):
    case = fuzz_model(seed)
    kw = dict(tu=case["tu"], tb=case["tb"], tc=case["tc"], td=[], **cases.sim_kwargs(case))
    p = util.write_fasta(case, workdir, "xsfuzz")
    o = po.oracle_simulate(p, want_trace=True, **kw)
    # the unmodified reference
    if po.have_ref():
        ctx = mp.get_context("fork")
        rx, tx = ctx.Pipe(duplex=False)
        pr = ctx.Process(target=_reference_child, args=(tx, p, kw))
        pr.start()
        tx.close()
        try:
            got = rx.recv() if rx.poll(60) else None
        except EOFError:
            got = None
        pr.join(10)
        if pr.is_alive():
            pr.kill()
            pytest.fail("the reference hangs on this model")
        if got is None:
            assert o.rc != 0, "the reference aborted (assert) where the oracle runs"
        else:
            assert got[0] == o.rc
            if o.rc == 0:
                assert got[1] == o.streams
                assert got[2] == tuple(o.cg)
    # the product's host side
    hs = util.HostSide(case)
    try:
        rc, cg, draws, nreads = hs.plan(cases.sim_kwargs(case))
        valid = all(hs.L.qthost_cat_valid(hs.h, c) for c in range(4))  # (qtipsim_run returns QTIPSIM_EXSCRIPT otherwise)
        assert (rc == 0 and valid) == (o.rc == 0)
        if o.rc == 0:
            assert cg == tuple(o.cg) and draws == o.draws and list(nreads) == list(o.wrote)
            wins = hs.windows(case["chunksz"], int(hs.L.qthost_olap(hs.h)))
            for cat in range(3):
                want = [(t[1], wins[t[2]][0], wins[t[2]][1] + t[3]) for t in o.trace if t[0] == cat]
                assert [(q[0], q[1], q[2]) for q in hs.plan_recs(cat, nreads[cat])] == want, cat
    finally:
        hs.close()
