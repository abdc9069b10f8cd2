"""Malformed CIGAR / MD:Z pairs (no GPU): the oracle's restatement of the SAM-fields -> edit-transcript path against the
unmodified reference (parse_cigar, mdz_to_list, cigar_to_edit_xscript, cigar_and_mdz_to_edit_xscript, lpos, rpos,
qtip_parse.cpp:164-186, :223-401; oracle/_ref/libreflabel.so), and the product's builder (xscript.h, the code
k_xs_lengths / k_xs_fill run per alignment, built for the host) against the oracle.

Inputs are well-formed alignments (test_oracle_pin.xscript_cases) with one or two random injuries: an op letter or a digit
of the CIGAR replaced, a byte dropped or doubled, MD:Z bytes replaced / dropped / inserted, a zero run added at either
end, the MD:Z tag removed, or either field replaced by a short oddity ("", "*", "M", "0M", "5M0I5M", "10M3N10M", "^A",
"5^^A5", ...).  The reference runs in forked children, because a third of these inputs end in one of its asserts:

* where the reference asserts or throws, the oracle must return a non-zero status (the record is refused);
* where it returns, transcript, clips and both positions are equal;
* ONE class is undefined in the reference and excluded: a transcript with at most one op after its leading soft clip
  (a one-column alignment, an all-S or an empty transcript).  rpos() (qtip_parse.cpp:173-186) steps over the leading S run
  AND the op after it before its assert(*xs != 0): with one op left the assert fires, with none the loop has already read
  past the terminator.  The oracle and the product return the transcript with rpos = pos + span - 1."""
import ctypes as C
import multiprocessing as mp
import os
import random

import pytest

import util
from oracle import pyoracle as po
from test_oracle_pin import xscript_cases


def injure(rnd, cigar, mdz):
    c, m = bytearray(cigar), bytearray(mdz) if mdz is not None else None
    k = rnd.random()
    if k < 0.15 and len(c) > 1:
        idx = [i for i, ch in enumerate(c) if chr(ch).isalpha() or chr(ch) == "="] or [0]
        c[rnd.choice(idx)] = ord(rnd.choice("MIDNSHP=XB"))
    elif k < 0.3 and len(c) > 1:
        idx = [i for i, ch in enumerate(c) if chr(ch).isdigit()] or [0]
        c[rnd.choice(idx)] = ord(rnd.choice("0123456789"))
    elif k < 0.4 and len(c) > 0:
        i = rnd.randrange(len(c))
        if rnd.random() < 0.5:
            del c[i]
        else:
            c.insert(i, c[i])
    elif k < 0.7 and m:
        i = rnd.randrange(len(m))
        r = rnd.random()
        if r < 0.25:
            m[i] = ord(rnd.choice("0123456789ACGTN^"))
        elif r < 0.5:
            del m[i]
        elif r < 0.75:
            m.insert(i, ord(rnd.choice("0ACGT^1")))
        else:
            m = bytearray(b"0") + m if rnd.random() < 0.5 else m + bytearray(b"0")
    elif k < 0.8:
        m = None
    elif k < 0.9:
        c = bytearray(rnd.choice([b"", b"*", b"10", b"M", b"0M", b"5M0I5M", b"3S", b"2H3S10M4S1H", b"10M3N10M"]))
    elif m is not None:
        m = bytearray(rnd.choice([b"", b"0", b"^A", b"10^", b"5^^A5", b"A", b"3A^A3", b"0A0"]))
    return bytes(c), (bytes(m) if m is not None else None)


def fuzz_cases(seed, n):
    rnd = random.Random(seed)
    base = xscript_cases(seed=seed + 50, n=300)
    out = []
    for _ in range(n):
        cigar, mdz, pos, _ = rnd.choice(base)
        if rnd.random() < 0.85:
            cigar, mdz = injure(rnd, cigar, mdz)
        if rnd.random() < 0.3:
            cigar, mdz = injure(rnd, cigar, mdz)
        out.append((cigar, mdz, pos))
    return out


def _reference_child(conn, recs, start):
    import faulthandler
    faulthandler.disable()  # the reference's asserts are expected outcomes here
    os.dup2(os.open(os.devnull, os.O_WRONLY), 2)
    r = po.ref_label()
    for i in range(start, len(recs)):
        conn.send((i, None))
        conn.send((i, po.edit_xscript(r.ref_edit_xscript, *recs[i])))
    conn.send((-1, None))


def reference_outcomes(recs):
    """Per record: the reference's (status, transcript, clips, lpos, rpos), or "abort" where it asserted / threw."""
    out, start = [None] * len(recs), 0
    ctx = mp.get_context("fork")
    while start < len(recs):
        rx, tx = ctx.Pipe(duplex=False)
        pr = ctx.Process(target=_reference_child, args=(tx, recs, start))
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
                    out[i] = res
        except EOFError:
            out[cur] = "abort"
        pr.join(10)
        start = cur + 1
    return out


def _norm(r):
    return r if r[0] == 0 else (r[0], b"", 0, 0, 0, 0)  # (the outputs of a refused record are unspecified)


@pytest.mark.skipif(not os.path.exists(po.REF_LABEL_SO), reason="oracle/_ref/libreflabel.so not built (needs /root/reference)")
@pytest.mark.parametrize("seed", [1, 3])
def test_injured_sam_fields(seed):
    recs = fuzz_cases(seed, 1000)
    ref = reference_outcomes(recs)
    o, L = po.orc(), util.hosttest()
    n_ok = n_abort = n_degenerate = 0
    for (cigar, mdz, pos), rr in zip(recs, ref):
        a = _norm(po.edit_xscript(o.orc_edit_xscript, cigar, mdz, pos))
        out = C.create_string_buffer(1 << 16)
        lc, rc_, lp, rp = C.c_int(), C.c_int(), C.c_uint64(), C.c_uint64()
        st = L.qthost_edit_xscript(cigar, mdz, pos, out, 1 << 16, C.byref(lc), C.byref(rc_), C.byref(lp), C.byref(rp))
        assert _norm((st, out.value, lc.value, rc_.value, lp.value, rp.value)) == a, (cigar, mdz, pos)  # product == oracle
        if a[0] == 0 and len(a[1].lstrip(b"S")) <= 1:
            n_degenerate += 1  # rpos() is undefined in the reference (see the module text)
            continue
        if rr == "abort":
            assert a[0] != 0, ("the reference asserts, the oracle accepts", cigar, mdz, pos, a)
            n_abort += 1
        else:
            assert _norm(rr) == a, (cigar, mdz, pos)
            n_ok += rr[0] == 0
    assert n_ok > 450 and n_abort > 200 and n_degenerate > 0
