"""Randomised malformed-but-legal FASTA text (no GPU): the oracle against the unmodified reference, and the product's host
logic (header resolution, 2-bit pack, window table, replay chain walk) against the oracle.

What the generator mixes: bases before the first header, empty records (header directly followed by a header), empty and
one-character names, names with blanks / tabs / '>', LF and CRLF, blank lines, blanks and tabs inside and after sequence
lines, a '>' in the middle of a sequence line, lower case, IUPAC codes and other bytes (`dna_upper`, fasta.cpp:18-35), lines of 1 to 10 000 characters, records of
0, 1, olap - 1, olap, olap + 1 and chunksz +- 1 bases, a missing final newline, trailing blank lines, windows from
olap + 1 bytes up.

One class of input is UNDEFINED in the reference and kept out of the reference comparison: a file that ends inside a header
line (no newline after the last '>').  The reference's header loop (fasta.cpp:54-71) never sees its '\\n', appends EOF to
the name until memory runs out and dies with std::bad_alloc.  Here such a file ends with an empty record
(test_header_cut_off_by_eof)."""
import ctypes as C
import multiprocessing as mp
import os
import random

import numpy as np
import pytest

import cases
import util
from oracle import pyoracle as po
from qtip_b200 import synth

SEEDS = list(range(60))


def weird_fasta(rnd, olap, chunksz, eof_in_header=False):
    parts = []
    if rnd.random() < 0.3:
        parts.append(rnd.choice([b"ACGTTGCA" * rnd.randint(1, 60) + b"\n", b"\n\n", b"  \t\n", b"acgtn\r\n" * 30]))
    for r in range(rnd.randint(1, 6)):
        name = rnd.choice([b"r%d" % r, b"", b"chr%d desc here" % r, b"x\ty", b" lead", b"a>b", b"n" * rnd.randint(1, 60)])
        nl = rnd.choice([b"\n", b"\n", b"\r\n"])
        n = rnd.choice([0, 0, 1, 2, olap - 1, olap, olap + 1, chunksz - 1, chunksz, chunksz + 1, 2 * chunksz - olap,
                        rnd.randint(1, 3000), rnd.randint(500, 6000)])
        alpha = rnd.choice([b"ACGT", b"ACGT", b"ACGTN", b"acgtACGT", b"ACGTRYKMSWBDHVN-.*", b"ACGTNNNNNNNN"])
        s = bytes(rnd.choice(alpha) for _ in range(n))
        width = rnd.choice([1, 7, 60, 70, 80, 10000])
        body = b""
        for i in range(0, len(s), width):
            ln = s[i:i + width]
            if rnd.random() < 0.05:
                body += nl  # blank line
            if rnd.random() < 0.05:
                ln = ln[:len(ln) // 2] + rnd.choice([b" ", b"\t", b"  "]) + ln[len(ln) // 2:]
            if rnd.random() < 0.02:
                ln = ln + b" "
            if rnd.random() < 0.02 and len(ln) > 2:
                ln = ln[:len(ln) // 3] + b">" + ln[len(ln) // 3:]  # a '>' in the middle of a sequence line starts a header
            body += ln + nl
        parts.append(b">" + name + nl + body)
    text = b"".join(parts)
    if rnd.random() < 0.3 and text.endswith(b"\n") and not text.endswith(b">\n"):
        last_hdr = text.rfind(b">")
        if text.find(b"\n", last_hdr) != len(text) - 1:  # (the newline dropped is not the last header's)
            text = text[:-1]
    if rnd.random() < 0.1:
        text += b"\n\n\n"
    if eof_in_header:
        text += rnd.choice([b">", b">cut off", b">a>b", b"> lead"])
    return text


def fuzz_case(seed, eof_in_header=False):
    rnd = random.Random(seed)
    L = rnd.choice([20, 50])
    chunksz = rnd.choice([L + 1, 64, 128, 512, 4096])
    text = weird_fasta(rnd, L, chunksz, eof_in_header)
    tu = synth.templates_unpaired(40, L, seed, max_x=3)
    return dict(fasta=np.frombuffer(text, dtype=np.uint8).copy(), tu=tu, tb=[], tc=[], td=[], chunksz=chunksz,
                kw=dict(fraction=200.0, function=3, min_u=1, min_c=1, min_d=1, min_b=1, cg=cases.cli_seed(seed % 1000 + 1)))


def check_host_against_oracle(case, workdir, name):
    """Records (name, length, bases), window table, and -- where the run is defined -- reads per category, draws
    consumed, generator state and every read's (template, record, offset)."""
    kw = cases.sim_kwargs(case)
    p = util.write_fasta(case, workdir, name)
    lib = po.orc()
    hs = util.HostSide(case)
    try:
        fa = lib.orc_fasta_load(p.encode())
        recs = hs.records()
        assert len(recs) == lib.orc_fasta_nrecords(fa)
        for i, (nm, ln, gs) in enumerate(recs):
            n = C.c_uint64()
            ptr = lib.orc_fasta_seq(fa, i, C.byref(n))
            assert nm == lib.orc_fasta_name(fa, i)
            assert ln == n.value
            assert hs.unpack(gs, ln) == C.string_at(ptr, ln)
        olap = int(hs.L.qthost_olap(hs.h))
        nw = C.c_size_t()
        ch = lib.orc_fasta_chunks(fa, kw["chunksz"], olap, C.byref(nw))
        want = [(ch[i].record, ch[i].refoff, ch[i].retsz, ch[i].n_count) for i in range(nw.value)]
        wins = hs.windows(kw["chunksz"], olap)
        assert wins == want
        lib.orc_fasta_free(fa)
        r = po.oracle_simulate(p, tu=case["tu"], want_trace=True, **kw)
        rc, cg, draws, nreads = hs.plan(kw)
        assert (rc == 0) == (r.rc == 0), (rc, r.rc, hs.L.qthost_error(hs.h))
        if r.rc == 0:
            assert cg == tuple(r.cg) and draws == r.draws and list(nreads) == list(r.wrote)
            got = [(q[0], q[1], q[2]) for q in hs.plan_recs(0, nreads[0])]
            assert got == [(t[1], wins[t[2]][0], wins[t[2]][1] + t[3]) for t in r.trace if t[0] == 0]
        return r
    finally:
        hs.close()


@pytest.mark.parametrize("seed", SEEDS)
def test_weird_fasta_host_logic_matches_oracle(seed, workdir):
    check_host_against_oracle(fuzz_case(seed), workdir, f"fuzz{seed}")


def _reference_child(conn, path, tu, kw):
    import faulthandler
    faulthandler.disable()  # an assert in the reference is an expected outcome here, not a crash to report
    devnull = os.open(os.devnull, os.O_WRONLY)
    os.dup2(devnull, 2)  # the "Wrote ..." lines and assertion messages
    r = po.ref_simulate(path, tu=tu, **kw)
    conn.send((r.rc, r.streams, tuple(r.cg)))
    conn.close()


@pytest.mark.skipif(not po.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("seed", SEEDS[:40])
def test_weird_fasta_oracle_matches_reference(seed, workdir):
    """The unmodified reference runs in a forked child: its assert(nslots > 0) (simplesim.cpp:252) aborts the process, and
    the oracle must refuse the same inputs; everywhere else the seven streams and the exit state are equal."""
    case = fuzz_case(seed)
    kw = cases.sim_kwargs(case)
    p = util.write_fasta(case, workdir, f"fuzz{seed}")
    b = po.oracle_simulate(p, tu=case["tu"], **kw)
    ctx = mp.get_context("fork")
    rx, tx = ctx.Pipe(duplex=False)
    pr = ctx.Process(target=_reference_child, args=(tx, p, case["tu"], kw))
    pr.start()
    tx.close()
    try:
        got = rx.recv() if rx.poll(60) else None
    except EOFError:  # the child died before it could answer
        got = None
    pr.join(10)
    if pr.is_alive():
        pr.kill()
        pytest.fail("the reference hangs on this input")
    if got is None:
        assert pr.exitcode != 0
        assert b.rc != 0, "the reference aborted (assert) where the oracle runs"
        return
    rc, streams, cg = got
    assert rc == b.rc == 0
    assert streams == b.streams
    assert cg == tuple(b.cg)


@pytest.mark.parametrize("seed", SEEDS[:12])
def test_header_cut_off_by_eof(seed, workdir):
    """A file that ends inside a header line: undefined in the reference (it exhausts memory, see the module text); the
    product and the oracle agree that it ends with an empty record of that name."""
    case = fuzz_case(seed, eof_in_header=True)
    check_host_against_oracle(case, workdir, f"fuzzeof{seed}")
    hs = util.HostSide(case)
    try:
        assert hs.records()[-1][1] == 0
    finally:
        hs.close()


@pytest.mark.gpu
@pytest.mark.parametrize("eof_in_header", [False, True])
@pytest.mark.parametrize("seed", SEEDS[:16])
def test_weird_fasta_device_path_bit_exact(seed, eof_in_header, workdir):
    """The same inputs through the C ABI on the GPU -- file loader, byte loader and the streaming path (K0 device pack,
    header resolution in api.cu, K2, K3): the seven streams and the exit state equal the oracle's, the inputs the
    reference refuses (assert(nslots > 0)) are refused with an error code."""
    case = fuzz_case(seed, eof_in_header)
    name = f"fuzzgpu{int(eof_in_header)}_{seed}"
    o = po.oracle_simulate(util.write_fasta(case, workdir, name), tu=case["tu"], **cases.sim_kwargs(case))
    for via in ("path", "bytes", "stream"):
        try:
            if via == "stream":
                streams, st, p, again = util.run_cuda_streaming(case)
                assert again == streams
            else:
                streams, st, p = util.run_cuda(case, workdir, name, via=via)
        except RuntimeError:
            assert o.rc != 0, via
            continue
        assert o.rc == 0, via
        assert streams == o.streams, via
        assert (p.cg1, p.cg2) == tuple(o.cg), via
