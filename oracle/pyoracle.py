"""TEST INFRASTRUCTURE ONLY: ctypes access to the CPU oracle (oracle/liboracle.so, our plain-C
restatement) and to the unmodified reference (oracle/_ref/libref.so, oracle/_ref/qtip-parse).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product (qtip_b200/, src/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import tempfile
from typing import List, Optional, Sequence, Tuple

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "liboracle.so")
REF_DIR = os.path.join(HERE, "_ref")
REF_SO = os.path.join(REF_DIR, "libref.so")
REF_PARSE = os.path.join(REF_DIR, "qtip-parse")
REF_LABEL_SO = os.path.join(REF_DIR, "libreflabel.so")

import sys

sys.path.insert(0, os.path.dirname(HERE))
from qtip_b200.templates import CTmplPair, CTmplUnp, marshal_paired, marshal_unpaired  # noqa: E402

OUT_NAMES = ("u", "b_1", "b_2", "c_1", "c_2", "d_1", "d_2")


class OrcRng(C.Structure):
    _fields_ = [("cg1", C.c_int32), ("cg2", C.c_int32), ("ndraws", C.c_uint64)]


class OrcChunk(C.Structure):
    _fields_ = [("record", C.c_uint32), ("refoff", C.c_uint64), ("retsz", C.c_uint64), ("n_count", C.c_uint64)]


class OrcTrace(C.Structure):
    _fields_ = [("cat", C.c_uint8), ("tmpl", C.c_uint32), ("chunk", C.c_uint32), ("off", C.c_uint64),
                ("draw_idx", C.c_uint64)]


class OrcOut(C.Structure):
    _fields_ = [("data", C.c_void_p * 7), ("len", C.c_size_t * 7), ("cap", C.c_size_t * 7),
                ("wrote", C.c_uint64 * 4), ("target", C.c_uint64 * 4), ("trace", C.POINTER(OrcTrace)),
                ("ntrace", C.c_size_t), ("captrace", C.c_size_t), ("want_trace", C.c_int)]


def build(ref: bool = True) -> None:
    """make -C oracle (the restatement; plus _ref/ when /root/reference is present)."""
    subprocess.run(["make", "-s", "-C", HERE, "oracle"] + (["ref"] if ref else []), check=True)


_orc = None
_ref = None


def orc() -> C.CDLL:
    global _orc
    if _orc is None:
        if not os.path.exists(ORACLE_SO):
            build(ref=False)
        lib = C.CDLL(ORACLE_SO)
        lib.orc_rng_initialize.argtypes = [C.POINTER(OrcRng)]
        lib.orc_rng_set_seed.argtypes = [C.POINTER(OrcRng), C.c_int32, C.c_int32]
        lib.orc_i4_uni.argtypes = [C.POINTER(OrcRng)]
        lib.orc_i4_uni.restype = C.c_int32
        lib.orc_r4_uni_01.argtypes = [C.POINTER(OrcRng)]
        lib.orc_r4_uni_01.restype = C.c_float
        lib.orc_ignbin.argtypes = [C.POINTER(OrcRng), C.c_int, C.c_float]
        lib.orc_ignbin.restype = C.c_int
        lib.orc_fasta_load.argtypes = [C.c_char_p]
        lib.orc_fasta_load.restype = C.c_void_p
        lib.orc_fasta_free.argtypes = [C.c_void_p]
        lib.orc_fasta_nrecords.argtypes = [C.c_void_p]
        lib.orc_fasta_nrecords.restype = C.c_size_t
        lib.orc_fasta_name.argtypes = [C.c_void_p, C.c_size_t]
        lib.orc_fasta_name.restype = C.c_char_p
        lib.orc_fasta_name_full.argtypes = [C.c_void_p, C.c_size_t]
        lib.orc_fasta_name_full.restype = C.c_char_p
        lib.orc_fasta_seq.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_uint64)]
        lib.orc_fasta_seq.restype = C.c_void_p
        lib.orc_fasta_filesize.argtypes = [C.c_void_p]
        lib.orc_fasta_filesize.restype = C.c_uint64
        lib.orc_fasta_chunks.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.POINTER(C.c_size_t)]
        lib.orc_fasta_chunks.restype = C.POINTER(OrcChunk)
        lib.orc_out_init.argtypes = [C.POINTER(OrcOut), C.c_int]
        lib.orc_out_free.argtypes = [C.POINTER(OrcOut)]
        lib.orc_apply_function.argtypes = [C.c_float, C.c_int, C.c_uint64, C.c_uint64]
        lib.orc_apply_function.restype = C.c_uint64
        lib.orc_xscript_rflen.argtypes = [C.c_char_p]
        lib.orc_xscript_rflen.restype = C.c_uint64
        lib.orc_simulate.argtypes = [
            C.c_void_p, C.c_uint64,
            C.POINTER(CTmplUnp), C.c_size_t, C.c_uint64,
            C.POINTER(CTmplUnp), C.c_size_t, C.c_uint64,
            C.POINTER(CTmplPair), C.c_size_t, C.c_uint64,
            C.POINTER(CTmplPair), C.c_size_t, C.c_uint64,
            C.c_float, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64,
            C.POINTER(OrcRng), C.POINTER(OrcOut)]
        lib.orc_simulate.restype = C.c_int
        lib.orc_set_correctness.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_uint64, C.c_int]
        lib.orc_set_correctness.restype = C.c_int
        lib.orc_edit_xscript.argtypes = [C.c_char_p, C.c_char_p, C.c_uint64, C.c_char_p, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                         C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        lib.orc_edit_xscript.restype = C.c_int
        _orc = lib
    return _orc


def have_ref() -> bool:
    return os.path.exists(REF_SO) and os.path.exists(REF_PARSE)


def ref() -> C.CDLL:
    global _ref
    if _ref is None:
        lib = C.CDLL(REF_SO)
        lib.ref_set_seed.argtypes = [C.c_int, C.c_int]
        lib.ref_get_state.argtypes = [C.POINTER(C.c_int), C.POINTER(C.c_int)]
        lib.ref_i4_uni.restype = C.c_int
        lib.ref_r4_uni_01.restype = C.c_float
        lib.ref_ignbin.argtypes = [C.c_int, C.c_float]
        lib.ref_ignbin.restype = C.c_int
        lib.ref_simulate.argtypes = [
            C.c_char_p, C.c_uint64,
            C.POINTER(CTmplUnp), C.c_uint64, C.c_uint64,
            C.POINTER(CTmplUnp), C.c_uint64, C.c_uint64,
            C.POINTER(CTmplPair), C.c_uint64, C.c_uint64,
            C.POINTER(CTmplPair), C.c_uint64, C.c_uint64,
            C.c_float, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64,
            C.POINTER(C.c_char_p)]
        lib.ref_simulate.restype = C.c_int
        lib.ref_initialize()
        _ref = lib
    return _ref


_reflabel = None


def ref_label() -> C.CDLL:
    """The unmodified reference's Alignment::set_correctness (oracle/ref_label_wrap.cpp -> _ref/libreflabel.so)."""
    global _reflabel
    if _reflabel is None:
        lib = C.CDLL(REF_LABEL_SO)
        lib.ref_set_correctness.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_uint64, C.c_int]
        lib.ref_set_correctness.restype = C.c_int
        lib.ref_edit_xscript.argtypes = [C.c_char_p, C.c_char_p, C.c_uint64, C.c_char_p, C.c_uint64, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                         C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        lib.ref_edit_xscript.restype = C.c_int
        _reflabel = lib
    return _reflabel


def edit_xscript(lib_fn, cigar: bytes, mdz, pos: int, cap: int = 1 << 16):
    """Calls orc_edit_xscript / ref_edit_xscript; returns (status, transcript, left_clip, right_clip, lpos, rpos)."""
    out = C.create_string_buffer(cap)
    lc, rc_ = C.c_int(), C.c_int()
    lp, rp = C.c_uint64(), C.c_uint64()
    st = lib_fn(cigar, mdz, pos, out, cap, C.byref(lc), C.byref(rc_), C.byref(lp), C.byref(rp))
    return (st, out.value if st == 0 else b"", lc.value, rc_.value, lp.value if st == 0 else 0, rp.value if st == 0 else 0)


class SimResult:
    def __init__(self, streams: List[bytes], wrote, target, cg, trace=None, rc=0, draws=0):
        self.streams = streams
        self.wrote = list(wrote)
        self.target = list(target)
        self.cg = cg  # generator state at exit
        self.trace = trace
        self.rc = rc
        self.draws = draws


def oracle_simulate(fasta_path: str, tu=(), tb=(), tc=(), td=(), added=None, chunksz=128 * 1024, fraction=30.0,
                    function=2, min_u=30000, min_c=30000, min_d=10000, min_b=10000, cg=(1234567890, 123456789),
                    want_trace=False) -> SimResult:
    """Runs our C restatement (orc_simulate).  `added` = (u, b, c, d) counts seen; defaults to kept."""
    lib = orc()
    fa = lib.orc_fasta_load(fasta_path.encode())
    if not fa:
        raise IOError(fasta_path)
    au, ku = marshal_unpaired(tu)
    ab, kb = marshal_unpaired(tb)
    ac, kc = marshal_paired(tc)
    ad, kd = marshal_paired(td)
    if added is None:
        added = (len(tu), len(tb), len(tc), len(td))
    rng = OrcRng(cg[0], cg[1], 0)
    out = OrcOut()
    lib.orc_out_init(C.byref(out), 1 if want_trace else 0)
    rc = lib.orc_simulate(fa, chunksz, au, len(tu), added[0], ab, len(tb), added[1], ac, len(tc), added[2], ad, len(td),
                          added[3], fraction, function, min_u, min_c, min_d, min_b, C.byref(rng), C.byref(out))
    streams = [C.string_at(out.data[i], out.len[i]) if out.len[i] else b"" for i in range(7)]
    trace = None
    if want_trace:
        trace = [(out.trace[i].cat, out.trace[i].tmpl, out.trace[i].chunk, out.trace[i].off, out.trace[i].draw_idx)
                 for i in range(out.ntrace)]
    res = SimResult(streams, out.wrote, out.target, (rng.cg1, rng.cg2), trace, rc, rng.ndraws)
    lib.orc_out_free(C.byref(out))
    lib.orc_fasta_free(fa)
    return res


def ref_simulate(fasta_path: str, tu=(), tb=(), tc=(), td=(), added=None, chunksz=128 * 1024, fraction=30.0,
                 function=2, min_u=30000, min_c=30000, min_d=10000, min_b=10000, cg=(1234567890, 123456789),
                 outdir: Optional[str] = None, read_back: bool = True) -> SimResult:
    """Runs the UNMODIFIED reference simulator (StreamingSimulator via oracle/_ref/libref.so)."""
    lib = ref()
    au, ku = marshal_unpaired(tu)
    ab, kb = marshal_unpaired(tb)
    ac, kc = marshal_paired(tc)
    ad, kd = marshal_paired(td)
    if added is None:
        added = (len(tu), len(tb), len(tc), len(td))
    tmp = None
    if outdir is None:
        tmp = tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
        outdir = tmp.name
    paths = [os.path.join(outdir, f"ref_reads_{n}.fastq").encode() for n in OUT_NAMES]
    arr = (C.c_char_p * 7)(*paths)
    lib.ref_set_seed(cg[0], cg[1])
    import time as _time
    _t0 = _time.perf_counter()
    rc = lib.ref_simulate(fasta_path.encode(), chunksz, au, len(tu), added[0], ab, len(tb), added[1], ac, len(tc),
                          added[2], ad, len(td), added[3], fraction, function, min_u, min_c, min_d, min_b, arr)
    _dt = _time.perf_counter() - _t0
    a, b = C.c_int(), C.c_int()
    lib.ref_get_state(C.byref(a), C.byref(b))
    streams = []
    if read_back:
        for p in paths:
            with open(p, "rb") as fh:
                streams.append(fh.read())
    else:
        streams = [os.path.getsize(p) for p in paths]
    if tmp is not None:
        tmp.cleanup()
    wrote = [0, 0, 0, 0]
    res = SimResult(streams, wrote, [0, 0, 0, 0], (a.value, b.value), None, rc)
    res.seconds = _dt  # time inside the reference's StreamingSimulator ctor + simulate_batch (+ template EList build)
    return res


def ref_qtip_parse(sam: str, fasta: str, outdir: str, seed: int, opts: Optional[dict] = None, exe: str = REF_PARSE,
                   mode: str = "s") -> Tuple[List[str], subprocess.CompletedProcess]:
    """Runs `qtip-parse` (reference CLI, qtip_parse.cpp:1594-1712) in simulate mode; returns the
    seven FASTQ paths."""
    o = {"seed": seed}
    o.update(opts or {})
    kv = []
    for k, v in o.items():
        kv += [str(k), str(v)]
    cmd = [exe, mode, "--"] + kv + ["--", sam, "--", fasta, "--", os.path.join(outdir, "rec"), "--",
                                   os.path.join(outdir, "tan")]
    cp = subprocess.run(cmd, capture_output=True, text=True)
    return [os.path.join(outdir, f"tan_reads_{n}.fastq") for n in OUT_NAMES], cp
