"""Shared helpers for the test-suite."""
from __future__ import annotations

import ctypes as C
import hashlib
import json
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "golden.json")
HOSTTEST_SO = os.path.join(ROOT, "qtip_b200", "libqtipsim_hosttest.so")


def golden():
    with open(GOLDEN) as fh:
        return json.load(fh)


def sha(b) -> str:
    return hashlib.sha256(bytes(b)).hexdigest()


def write_fasta(case, workdir, name) -> str:
    p = os.path.join(workdir, f"{name}.fa")
    if not os.path.exists(p):
        case["fasta"].tofile(p)
    return p


_host = None


def hosttest() -> C.CDLL:
    """CPU-only hooks over libqtipsim's host logic (qtip_b200/csrc/hosttest.cpp)."""
    global _host
    if _host is None:
        if not os.path.exists(HOSTTEST_SO):
            subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "qtip_b200", "csrc"), HOSTTEST_SO], check=True)
        L = C.CDLL(HOSTTEST_SO)
        L.qthost_create.restype = C.c_void_p
        L.qthost_destroy.argtypes = [C.c_void_p]
        L.qthost_error.argtypes = [C.c_void_p]
        L.qthost_error.restype = C.c_char_p
        L.qthost_parse.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
        L.qthost_nrecords.argtypes = [C.c_void_p]
        L.qthost_nrecords.restype = C.c_uint64
        L.qthost_record.argtypes = [C.c_void_p, C.c_uint64, C.c_char_p, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.qthost_record.restype = C.c_uint64
        L.qthost_unpack.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_char_p]
        L.qthost_set_unpaired.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_uint64]
        L.qthost_set_paired.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_uint64]
        L.qthost_olap.argtypes = [C.c_void_p]
        L.qthost_olap.restype = C.c_uint64
        L.qthost_cat_valid.argtypes = [C.c_void_p, C.c_int]
        L.qthost_windows.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64]
        L.qthost_windows.restype = C.c_uint64
        L.qthost_window.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.qthost_plan.argtypes = [C.c_void_p, C.c_uint64, C.c_float, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64,
                                  C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64 * 4),
                                  C.c_uint64, C.c_uint64]
        L.qthost_plan_rec.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32),
                                      C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.qthost_rng_step.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_int32), C.POINTER(C.c_float)]
        L.qthost_rng_jump.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_uint64]
        L.qthost_ignbin.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_int, C.c_float]
        L.qthost_ignbin.restype = C.c_int
        L.qthost_ignbin_tab.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_int, C.c_float]
        L.qthost_ignbin_tab.restype = C.c_int
        L.qthost_base_thresholds.argtypes = [C.POINTER(C.c_int32), C.c_int32, C.c_int32]
        L.qthost_base_thresholds.restype = C.c_uint64
        L.qthost_label.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_uint64, C.c_int]
        L.qthost_label.restype = C.c_int
        L.qthost_edit_xscript.argtypes = [C.c_char_p, C.c_char_p, C.c_uint64, C.c_char_p, C.c_uint32, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                          C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.qthost_edit_xscript.restype = C.c_int
        L.qthost_logf.argtypes = [C.c_float]
        L.qthost_logf.restype = C.c_float
        L.qthost_logf_sweep.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32]
        L.qthost_logf_sweep.restype = C.c_uint64
        _host = L
    return _host


class HostSide:
    """One hosttest handle loaded with a case."""

    def __init__(self, case):
        from qtip_b200.templates import marshal_paired, marshal_unpaired
        self.L = hosttest()
        self.h = C.c_void_p(self.L.qthost_create())
        fa = np.ascontiguousarray(case["fasta"])
        self.L.qthost_parse(self.h, C.c_void_p(fa.ctypes.data), fa.size)
        added = case.get("added") or (len(case["tu"]), len(case["tb"]), len(case["tc"]), len(case["td"]))
        self._keep = []
        for cat, key, m in ((0, "tu", marshal_unpaired), (1, "tb", marshal_unpaired)):
            arr, keep = m(case[key])
            self._keep.append((arr, keep))
            self.L.qthost_set_unpaired(self.h, cat, C.cast(arr, C.c_void_p), len(case[key]), added[cat])
        for cat, key in ((2, "tc"), (3, "td")):
            arr, keep = marshal_paired(case[key])
            self._keep.append((arr, keep))
            self.L.qthost_set_paired(self.h, cat, C.cast(arr, C.c_void_p), len(case[key]), added[cat])

    def close(self):
        if self.h:
            self.L.qthost_destroy(self.h)
            self.h = None

    def records(self):
        out = []
        for i in range(self.L.qthost_nrecords(self.h)):
            buf = C.create_string_buffer(4096)
            ln, gs = C.c_uint64(), C.c_uint64()
            self.L.qthost_record(self.h, i, buf, 4096, C.byref(ln), C.byref(gs))
            out.append((buf.value, ln.value, gs.value))
        return out

    def unpack(self, g0, n) -> bytes:
        buf = C.create_string_buffer(n)
        self.L.qthost_unpack(self.h, g0, n, buf)
        return buf.raw

    def windows(self, chunksz, olap):
        n = self.L.qthost_windows(self.h, chunksz, olap)
        out = []
        for i in range(n):
            r, ro, rs, nc = C.c_uint32(), C.c_uint64(), C.c_uint64(), C.c_uint64()
            self.L.qthost_window(self.h, i, C.byref(r), C.byref(ro), C.byref(rs), C.byref(nc))
            out.append((r.value, ro.value, rs.value, nc.value))
        return out

    def plan(self, kw, win_lo=0, win_hi=1 << 60, cg=None):
        cg = cg or kw["cg"]
        cg1, cg2 = C.c_int32(cg[0]), C.c_int32(cg[1])
        draws = C.c_uint64()
        nr = (C.c_uint64 * 4)()
        rc = self.L.qthost_plan(self.h, kw["chunksz"], kw["fraction"], kw["function"], kw["min_u"], kw["min_c"],
                                kw["min_d"], kw["min_b"], C.byref(cg1), C.byref(cg2), C.byref(draws), C.byref(nr),
                                win_lo, win_hi)
        return rc, (cg1.value, cg2.value), draws.value, list(nr)

    def plan_recs(self, cat, n):
        out = []
        for i in range(n):
            t, r, s1, s2 = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_uint32()
            ro = C.c_uint64()
            self.L.qthost_plan_rec(self.h, cat, i, C.byref(t), C.byref(r), C.byref(ro), C.byref(s1), C.byref(s2))
            out.append((t.value, r.value, ro.value, s1.value, s2.value))
        return out


def run_cuda(case, workdir, name, rng_mode=0, philox_seed=0, shard=(0, 1), via="path", cg=None):
    """Runs the case through libqtipsim's C ABI on the GPU; returns (streams, stats, params)."""
    from qtip_b200 import sim
    eng = sim.Engine()
    try:
        if via == "path":
            eng.load_fasta(write_fasta(case, workdir, name))
        else:
            eng.load_fasta_bytes(case["fasta"])
        added = case.get("added") or (len(case["tu"]), len(case["tb"]), len(case["tc"]), len(case["td"]))
        eng.set_unpaired(0, case["tu"], added[0])
        eng.set_unpaired(1, case["tb"], added[1])
        eng.set_paired(2, case["tc"], added[2])
        eng.set_paired(3, case["td"], added[3])
        kw = case["kw"]
        p = sim.make_params(case.get("chunksz", 128 * 1024), kw["fraction"], kw["function"], kw["min_u"], kw["min_c"],
                            kw["min_d"], kw["min_b"], rng_mode, cg=cg or kw["cg"], philox_seed=philox_seed,
                            shard_rank=shard[0], shard_count=shard[1])
        st = eng.run(p)
        streams = eng.results(st)
        return streams, st, p
    finally:
        eng.close()


def run_cuda_streaming(case, rng_mode=0, philox_seed=0, cg=None):
    """The case through the streaming end-to-end path (qtipsim_stream_begin / qtipsim_stream_run); returns
    (streams, stats, params) like run_cuda."""
    import numpy as np
    from qtip_b200 import sim
    eng = sim.Engine()
    try:
        fa = np.ascontiguousarray(np.frombuffer(bytes(case["fasta"]), dtype=np.uint8))
        eng.stream_begin(fa)
        added = case.get("added") or (len(case["tu"]), len(case["tb"]), len(case["tc"]), len(case["td"]))
        eng.set_unpaired(0, case["tu"], added[0])
        eng.set_unpaired(1, case["tb"], added[1])
        eng.set_paired(2, case["tc"], added[2])
        eng.set_paired(3, case["td"], added[3])
        kw = case["kw"]
        p = sim.make_params(case.get("chunksz", 128 * 1024), kw["fraction"], kw["function"], kw["min_u"], kw["min_c"],
                            kw["min_d"], kw["min_b"], rng_mode, cg=cg or kw["cg"], philox_seed=philox_seed)
        cap = 64 << 20
        bufs = [np.zeros(cap, dtype=np.uint8) for _ in range(7)]
        st = eng.stream_run(p, [b.ctypes.data for b in bufs], [cap] * 7)
        streams = [bufs[i][:int(st.bytes[i])].tobytes() for i in range(7)]
        # the genome is resident afterwards: a plain run on the same engine gives the same bytes again
        p2 = sim.make_params(case.get("chunksz", 128 * 1024), kw["fraction"], kw["function"], kw["min_u"], kw["min_c"],
                             kw["min_d"], kw["min_b"], rng_mode, cg=cg or kw["cg"], philox_seed=philox_seed)
        st2 = eng.run(p2)
        again = eng.results(st2)
        return streams, st, p, again
    finally:
        eng.close()
