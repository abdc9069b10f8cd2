"""Injured read names for the correctness-labelling operator (no GPU): simulator-format and wgsim-format names
(test_oracle_pin.label_cases) with a byte dropped or replaced, a truncation, separators or an overlong number inserted,
fields shuffled or removed, or a short oddity ("qtip!", "qtip!:a:+", "_", ...), and FLAGs that contradict the name.  The
unmodified reference's Alignment::set_correctness (qtip_parse.cpp:519-713, oracle/_ref/libreflabel.so) runs in forked
children: about one input in twenty ends in one of its asserts (an unpaired FLAG on a pair's name, a name that stops
inside a field); there the operator has no reference value and returns a label all the same.  Everywhere else the oracle
equals the reference, and the product's labelling function (label.h, what k_label_correctness runs per record, built
for the host) equals the oracle on every input."""
import multiprocessing as mp
import os
import random

import pytest

import util
from oracle import pyoracle as po
from test_oracle_pin import label_cases


def injure(rnd, q):
    b = bytearray(q)
    k = rnd.random()
    if k < 0.2 and len(b) > 1:
        del b[rnd.randrange(len(b))]
    elif k < 0.4 and len(b) > 0:
        b[rnd.randrange(len(b))] = ord(rnd.choice(":_+-!0123456789abuxcdq "))
    elif k < 0.55:
        b = b[:rnd.randrange(len(b) + 1)]
    elif k < 0.7:
        i = rnd.randrange(len(b) + 1)
        b[i:i] = rnd.choice([b":", b"::", b"_", b"-", b"+", b"99999999999999999999", b" "])
    elif k < 0.8:
        parts = bytes(b).split(b":")
        rnd.shuffle(parts)
        b = bytearray(b":".join(parts))
    elif k < 0.9:
        parts = bytes(b).split(b":")
        if len(parts) > 2:
            del parts[rnd.randrange(len(parts))]
        b = bytearray(b":".join(parts))
    else:
        b = bytearray(rnd.choice([b"", b"qtip!", b"qtip!:", b"qtip!:a", b"qtip!:a:+", b"qtip!:a:+:5", b"qtip!:a:+:5:3",
                                  b"qtip!:a:+:5:3:", b"qtip!:a:+:5:3:z", b"_", b"a_1"]))
    return bytes(b)


def _reference_child(conn, recs, start, wiggle):
    import faulthandler
    faulthandler.disable()  # the reference's asserts are expected outcomes here
    os.dup2(os.open(os.devnull, os.O_WRONLY), 2)
    r = po.ref_label()
    for i in range(start, len(recs)):
        conn.send((i, None))
        conn.send((i, r.ref_set_correctness(*recs[i], wiggle)))
    conn.send((-1, None))


@pytest.mark.skipif(not os.path.exists(po.REF_LABEL_SO), reason="oracle/_ref/libreflabel.so not built (needs /root/reference)")
@pytest.mark.parametrize("seed", [1, 2])
def test_injured_read_names(seed):
    rnd = random.Random(seed)
    base = label_cases(seed=seed, n=600)
    recs = []
    for _ in range(2000):
        q, rn, flag, pos = rnd.choice(base)
        if rnd.random() < 0.8:
            q = injure(rnd, q)
        if rnd.random() < 0.2:
            flag = rnd.choice([0, 16, 1, 65, 129, 81, 145, 4, 256, 2048 + 65, 3])
        recs.append((q, rn, flag, pos))
    wiggle = 30
    ref, start = [None] * len(recs), 0
    ctx = mp.get_context("fork")
    while start < len(recs):
        rx, tx = ctx.Pipe(duplex=False)
        pr = ctx.Process(target=_reference_child, args=(tx, recs, start, wiggle))
        pr.start()
        tx.close()
        cur = start
        try:
            while True:
                assert rx.poll(60), "the reference hangs"
                i, res = rx.recv()
                if i < 0:
                    cur = len(recs)
                    break
                if res is None:
                    cur = i
                else:
                    ref[i] = res
        except EOFError:
            ref[cur] = "abort"
        pr.join(10)
        start = cur + 1
    o, L = po.orc(), util.hosttest()
    seen, aborted = set(), 0
    for (q, rn, flag, pos), rr in zip(recs, ref):
        a = o.orc_set_correctness(q, rn, flag, pos, wiggle)
        assert L.qthost_label(q, rn, flag, pos, wiggle) == a, (q, rn, flag, pos)  # product == oracle, always
        assert a in (-1, 0, 1)
        if rr == "abort":
            aborted += 1
            continue
        assert rr == a, (q, rn, flag, pos, rr, a)
        seen.add(a)
    assert seen == {-1, 0, 1} and 20 < aborted < 400
