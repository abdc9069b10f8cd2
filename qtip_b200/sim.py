"""Host-side mirror of the reference's simulator interface, forwarding to libqtipsim.

Same names, argument order and error behaviour as the reference classes:

* ``InputModelUnpaired`` / ``InputModelPaired``  -- src/input_model.h:22-167
* ``StreamingSimulator``                         -- src/simplesim.h:203-299,
  ``simulate_batch``                               src/simplesim.cpp:195-407

The work happens on the GPU behind the C ABI (include/qtipsim.h); nothing here computes
reads on the CPU.
"""
from __future__ import annotations

import ctypes as C
import sys
from typing import List, Optional, Sequence

import numpy as np

from . import _abi, rnglib
from ._abi import CAT_B, CAT_C, CAT_D, CAT_U, FUNC_CONST, FUNC_LINEAR, FUNC_SQRT, NOUT, RNG_PHILOX, RNG_REPLAY
from .templates import TemplatePaired, TemplateUnpaired, marshal_paired, marshal_unpaired


class QtipSimError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libqtipsim error {code}: {msg}")
        self.code = code


class InputModelUnpaired:
    """input_model.h:22-94"""

    def __init__(self, ts: Sequence[TemplateUnpaired], n: int, fraction_even: float = 1.0, low_score_bias: float = 1.0):
        self.ts = list(ts)
        self.n = int(n)
        self.fraction_even = fraction_even  # unused, as in the reference
        self.low_score_bias = low_score_bias  # unused
        self._max = max((t.reflen() for t in self.ts), default=0)
        self._avg = float(np.float32(sum(np.float32(t.reflen()) / np.float32(len(self.ts)) for t in self.ts))) if self.ts else 0.0

    def empty(self) -> bool:
        return not self.ts

    def num_added(self) -> int:
        return self.n

    def avg_len(self) -> float:
        return self._avg

    def max_len(self) -> int:
        return self._max


class InputModelPaired:
    """input_model.h:96-167"""

    def __init__(self, ts: Sequence[TemplatePaired], n: int, fraction_even: float = 1.0, low_score_bias: float = 1.0):
        self.ts = list(ts)
        self.n = int(n)
        self.fraction_even = fraction_even
        self.low_score_bias = low_score_bias
        self._max = max((t.fraglen for t in self.ts), default=0)
        self._avg = float(sum(t.fraglen for t in self.ts) / len(self.ts)) if self.ts else 0.0

    def empty(self) -> bool:
        return not self.ts

    def num_added(self) -> int:
        return self.n

    def avg_len(self) -> float:
        return self._avg

    def max_len(self) -> int:
        return self._max


class Engine:
    """Thin RAII wrapper over one ``qtipsim_t`` handle."""

    def __init__(self, device: int = -1):
        self.lib = _abi.load()
        h = C.c_void_p()
        rc = self.lib.qtipsim_create(C.byref(h), device)
        if rc != 0:
            raise QtipSimError(rc, self.lib.qtipsim_last_error(None).decode())
        self.h = h
        self._keep = []
        self._marsh = {}

    def close(self):
        if getattr(self, "h", None):
            self.lib.qtipsim_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc: int):
        if rc != 0:
            raise QtipSimError(rc, self.lib.qtipsim_last_error(self.h).decode())

    def load_fasta(self, path: str):
        arr = (C.c_char_p * 1)(path.encode())
        self.check(self.lib.qtipsim_load_fasta(self.h, arr, 1))

    def load_fasta_paths(self, paths: Sequence[str]):
        arr = (C.c_char_p * len(paths))(*[p.encode() for p in paths])
        self.check(self.lib.qtipsim_load_fasta(self.h, arr, len(paths)))

    def load_fasta_bytes(self, data) -> None:
        """data: bytes, or a numpy uint8 array (e.g. pinned host memory)"""
        if isinstance(data, (bytes, bytearray)):
            buf = (C.c_char * len(data)).from_buffer_copy(data) if isinstance(data, bytes) else (C.c_char * len(data)).from_buffer(data)
            self.check(self.lib.qtipsim_load_fasta_bytes(self.h, C.cast(buf, C.c_void_p), len(data)))
        else:
            a = np.ascontiguousarray(data, dtype=np.uint8)
            self.check(self.lib.qtipsim_load_fasta_bytes(self.h, C.c_void_p(a.ctypes.data), a.size))

    def model_begin(self, k: Sequence[int]) -> None:
        """Four empty reservoirs of capacity k[cat] (input-model-size); qtipsim_model_begin."""
        self.check(self.lib.qtipsim_model_begin(self.h, (C.c_uint64 * 4)(*[int(x) for x in k])))

    def model_add(self, batch: "_abi.ModelBatch", cg) -> tuple:
        """Offers the batch's records (SAM order) to the reservoirs; cg = rnglib state in, returns the state out."""
        cg1, cg2 = C.c_int32(int(cg[0])), C.c_int32(int(cg[1]))
        self.check(self.lib.qtipsim_model_add(self.h, C.byref(batch), C.byref(cg1), C.byref(cg2)))
        return cg1.value, cg2.value

    def model_commit(self) -> tuple:
        """The kept templates become the four input models; returns (kept[4], added[4])."""
        kept, added = (C.c_uint64 * 4)(), (C.c_uint64 * 4)()
        self.check(self.lib.qtipsim_model_commit(self.h, kept, added))
        return list(kept), list(added)

    def stream_begin(self, data) -> None:
        """Streaming end-to-end path, step 1: starts the chunked upload of the FASTA text (a numpy uint8 array, pinned
        for asynchronous copies; it must stay alive until stream_run returns) and returns at once."""
        a = np.ascontiguousarray(data, dtype=np.uint8)
        self._stream_text = a
        self.check(self.lib.qtipsim_stream_begin(self.h, C.c_void_p(a.ctypes.data), a.size))

    def stream_run(self, params: _abi.RunParams, dst_ptrs: Sequence[int], caps: Sequence[int]) -> _abi.Stats:
        """Step 2 (templates set): packs and simulates record by record while later text is still uploading and earlier
        reads are downloading; appends every stream's FASTQ bytes to the host buffer at dst_ptrs[i] (caps[i] bytes).
        Returns the stats (bytes[i] = bytes written to buffer i).  Same bytes as load_fasta_bytes + run + result."""
        st = _abi.Stats()
        d = (C.c_void_p * _abi.NOUT)(*[C.c_void_p(int(x) if x else None) for x in dst_ptrs])
        c = (C.c_uint64 * _abi.NOUT)(*[int(x) for x in caps])
        try:
            self.check(self.lib.qtipsim_stream_run(self.h, C.byref(params), C.byref(st), d, c))
        finally:
            self._stream_text = None
        return st

    def load_fasta_bytes_async(self, data):
        """Starts load_fasta_bytes on a worker thread (ctypes releases the GIL during the call) so that the host
        can compile the templates while the FASTA text crosses PCIe and is packed; join with wait_genome()."""
        import threading
        box = {}

        def work():
            try:
                self.load_fasta_bytes(data)
            except BaseException as e:  # re-raised by wait_genome
                box["err"] = e

        th = threading.Thread(target=work, daemon=True)
        th.start()
        self._genome_job = (th, box)

    def wait_genome(self) -> None:
        job = getattr(self, "_genome_job", None)
        if job:
            job[0].join()
            self._genome_job = None
            if "err" in job[1]:
                raise job[1]["err"]

    def _marshalled(self, cat: int, ts, fn):
        # the ctypes array for a template list is built once and reused (the C call itself is cheap)
        key = (cat, id(ts), len(ts))
        hit = self._marsh.get(cat)
        if hit is None or hit[0] != key:
            arr, keep = fn(ts)
            self._marsh[cat] = (key, arr, keep)
        return self._marsh[cat][1]

    def set_unpaired(self, cat: int, ts: Sequence[TemplateUnpaired], added: Optional[int] = None):
        arr = self._marshalled(cat, ts, marshal_unpaired)
        self.check(self.lib.qtipsim_set_templates_unpaired(self.h, cat, arr, len(ts), len(ts) if added is None else added))

    def set_paired(self, cat: int, ts: Sequence[TemplatePaired], added: Optional[int] = None):
        arr = self._marshalled(cat, ts, marshal_paired)
        self.check(self.lib.qtipsim_set_templates_paired(self.h, cat, arr, len(ts), len(ts) if added is None else added))

    def olap(self) -> int:
        return int(self.lib.qtipsim_olap(self.h))

    def num_estimated_bases(self) -> int:
        return int(self.lib.qtipsim_num_estimated_bases(self.h))

    def run(self, params: _abi.RunParams) -> _abi.Stats:
        st = _abi.Stats()
        self.check(self.lib.qtipsim_run(self.h, C.byref(params), C.byref(st)))
        return st

    def result(self, stream: int, nbytes: int) -> np.ndarray:
        out = np.empty(nbytes, dtype=np.uint8)
        if nbytes:
            self.check(self.lib.qtipsim_result_copy(self.h, stream, C.c_void_p(out.ctypes.data), nbytes))
        return out

    def results(self, st: _abi.Stats) -> List[bytes]:
        return [self.result(s, int(st.bytes[s])).tobytes() for s in range(NOUT)]

    def stream_fds(self, fds: Sequence[int]):
        """All seven streams written concurrently (FIFO / pipe consumers that read paired files in lock step)."""
        arr = (C.c_int * NOUT)(*fds)
        self.check(self.lib.qtipsim_result_stream_fds(self.h, arr))

    def result_interleaved(self, cat: int, st: _abi.Stats) -> bytes:
        """Mate 1 / mate 2 records of a pair category as one interleaved FASTQ stream."""
        s0 = 2 * cat - 1
        cap = int(st.bytes[s0]) + int(st.bytes[s0 + 1])
        out = np.empty(max(cap, 1), dtype=np.uint8)
        n = C.c_uint64()
        self.check(self.lib.qtipsim_result_interleaved(self.h, cat, C.c_void_p(out.ctypes.data), cap, C.byref(n)))
        return out[:int(n.value)].tobytes()

    def load_fasta_cached(self, fasta_path: str, cache_path: str) -> bool:
        """Loads the genome from the packed cache if it is stamped with the FASTA's size and mtime, else packs the FASTA and
        writes the cache; returns True on a cache hit."""
        hit = C.c_int(0)
        self.check(self.lib.qtipsim_load_fasta_cached(self.h, fasta_path.encode(), cache_path.encode(), C.byref(hit)))
        return bool(hit.value)

    def genome_save(self, cache_path: str, fasta_path: str):
        self.check(self.lib.qtipsim_genome_save(self.h, cache_path.encode(), fasta_path.encode()))

    def label_correctness(self, qnames: Sequence[bytes], rnames: Sequence[bytes], flags: Sequence[int], pos: Sequence[int],
                          wiggle: int = 30) -> np.ndarray:
        """Alignment::set_correctness (qtip_parse.cpp:519-713) for a batch of aligned tandem-SAM records; returns int8
        labels (1 correct, 0 incorrect, -1 not a simulated-read name)."""
        n = len(qnames)
        out = np.zeros(n, dtype=np.int8)
        if not n:
            return out
        qa = b"".join(q + b"\0" for q in qnames)
        ra = b"".join(r + b"\0" for r in rnames)
        qo = np.zeros(n, dtype=np.uint64)
        ro = np.zeros(n, dtype=np.uint64)
        qo[1:] = np.cumsum([len(q) + 1 for q in qnames[:-1]], dtype=np.uint64)
        ro[1:] = np.cumsum([len(r) + 1 for r in rnames[:-1]], dtype=np.uint64)
        fl = np.ascontiguousarray(flags, dtype=np.int32)
        ps = np.ascontiguousarray(pos, dtype=np.uint64)
        P64, P32, P8 = C.POINTER(C.c_uint64), C.POINTER(C.c_int32), C.POINTER(C.c_int8)
        self.check(self.lib.qtipsim_label_correctness(self.h, qa, len(qa), qo.ctypes.data_as(P64), ra, len(ra), ro.ctypes.data_as(P64),
                                                      fl.ctypes.data_as(P32), ps.ctypes.data_as(P64), n, wiggle, out.ctypes.data_as(P8)))
        return out

    def edit_xscripts(self, cigars: Sequence[bytes], mdzs: Sequence[Optional[bytes]], pos: Sequence[int]):
        """SAM CIGAR (+ MD:Z) -> edit transcripts on the device (qtip_parse.cpp:223-401); returns a list of
        (status, transcript, left_clip, right_clip, lpos, rpos) per alignment."""
        n = len(cigars)
        if not n:
            return []
        ca = b"".join(c + b"\0" for c in cigars)
        ma = b"".join(m + b"\0" for m in mdzs if m is not None)
        co = np.zeros(n, dtype=np.uint64)
        co[1:] = np.cumsum([len(c) + 1 for c in cigars[:-1]], dtype=np.uint64)
        mo = np.full(n, np.iinfo(np.uint64).max, dtype=np.uint64)
        acc = 0
        for i, m in enumerate(mdzs):
            if m is not None:
                mo[i] = acc
                acc += len(m) + 1
        ps = np.ascontiguousarray(pos, dtype=np.uint64)
        xo = np.zeros(n + 1, dtype=np.uint64)
        lc, rc_, st = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, np.int32)
        lp, rp = np.zeros(n, np.uint64), np.zeros(n, np.uint64)
        P64, P32 = C.POINTER(C.c_uint64), C.POINTER(C.c_int32)
        cap = 1 << 16
        for _ in range(2):  # the second call has the size the first one reported
            buf = np.zeros(cap, dtype=np.uint8)
            rc = self.lib.qtipsim_edit_xscripts(self.h, ca, len(ca), co.ctypes.data_as(P64), ma if ma else None, len(ma), mo.ctypes.data_as(P64),
                                                ps.ctypes.data_as(P64), n, C.c_void_p(buf.ctypes.data), cap, xo.ctypes.data_as(P64),
                                                lc.ctypes.data_as(P32), rc_.ctypes.data_as(P32), lp.ctypes.data_as(P64), rp.ctypes.data_as(P64),
                                                st.ctypes.data_as(P32))
            if rc == -6 and int(xo[n]) > cap:
                cap = int(xo[n])
                continue
            self.check(rc)
            break
        raw = buf.tobytes()
        out = []
        for i in range(n):
            xs = raw[int(xo[i]):int(xo[i + 1]) - 1]
            ok = st[i] == 0
            out.append((int(st[i]), xs if ok else b"", int(lc[i]), int(rc_[i]), int(lp[i]) if ok else 0, int(rp[i]) if ok else 0))
        return out

    def result_device_ptr(self, stream: int):
        p = C.c_void_p()
        n = C.c_uint64()
        self.check(self.lib.qtipsim_result_device_ptr(self.h, stream, C.byref(p), C.byref(n)))
        return p.value, int(n.value)

    def write_fds(self, fds: Sequence[int]):
        arr = (C.c_int * NOUT)(*fds)
        self.check(self.lib.qtipsim_result_write_fds(self.h, arr))


def make_params(chunksz=128 * 1024, fraction=30.0, function=FUNC_SQRT, min_u=30000, min_c=30000, min_d=10000,
                min_b=10000, rng_mode=RNG_REPLAY, cg=None, philox_seed=0, shard_rank=0, shard_count=1) -> _abi.RunParams:
    """Defaults are qtip_parse.cpp:797-807 (C++ side) and :1840 (chunksz)."""
    p = _abi.RunParams()
    p.chunksz = chunksz
    p.fraction = fraction
    p.function = function
    p.min_u, p.min_c, p.min_d, p.min_b = min_u, min_c, min_d, min_b
    p.rng_mode = rng_mode
    cg1, cg2 = cg if cg is not None else rnglib.get_state()
    p.cg1, p.cg2 = cg1, cg2
    p.philox_seed = philox_seed
    p.shard_rank, p.shard_count = shard_rank, shard_count
    return p


class StreamingSimulator:
    """simplesim.h:203-299.  File arguments are binary file objects (the caller opens,
    buffers and closes them, as qtip_parse.cpp:1831-1866 does with its FILE*s)."""

    def __init__(self, fns: Sequence[str], chunksz: int, model_u: InputModelUnpaired, model_b: InputModelUnpaired,
                 model_c: InputModelPaired, model_d: InputModelPaired, fh_u, fh_b_1, fh_b_2, fh_c_1, fh_c_2, fh_d_1,
                 fh_d_2, device: int = -1, rng_mode: int = RNG_REPLAY, philox_seed: int = 0):
        self.chunksz = int(chunksz)
        self.models = (model_u, model_b, model_c, model_d)
        self.fhs = (fh_u, fh_b_1, fh_b_2, fh_c_1, fh_c_2, fh_d_1, fh_d_2)
        self.rng_mode = rng_mode
        self.philox_seed = philox_seed
        self.eng = Engine(device)
        # olap_ = max over models (simplesim.h:228-230); assert(chunksz > olap), fasta.h:41
        self.olap_ = max(m.max_len() for m in self.models)
        if not self.chunksz > self.olap_:
            raise AssertionError("chunksz > olap")
        self.eng.load_fasta_paths(list(fns))
        self.eng.set_unpaired(CAT_U, model_u.ts, model_u.num_added())
        self.eng.set_unpaired(CAT_B, model_b.ts, model_b.num_added())
        self.eng.set_paired(CAT_C, model_c.ts, model_c.num_added())
        self.eng.set_paired(CAT_D, model_d.ts, model_d.num_added())
        self.last_stats = None

    def num_estimated_bases(self) -> int:
        """simplesim.h:263-265"""
        return self.eng.num_estimated_bases()

    def simulate_batch(self, fraction: float, function: int, min_u: int, min_c: int, min_d: int, min_b: int):
        """simplesim.cpp:195-407 (argument order u, c, d, b as in the reference)."""
        p = make_params(self.chunksz, fraction, function, min_u, min_c, min_d, min_b, self.rng_mode,
                        philox_seed=self.philox_seed)
        st = self.eng.run(p)
        if self.rng_mode == RNG_REPLAY:
            rnglib.set_seed(p.cg1, p.cg2)  # the reference's global generator has advanced
        for s, fh in enumerate(self.fhs):
            if fh is not None and st.bytes[s]:
                fh.write(self.eng.result(s, int(st.bytes[s])).data)
        # simplesim.cpp:399-406
        names = ("unpaired tandem reads", "bad-end tandem reads", "concordant tandem pairs", "discordant tandem pairs")
        for i, nm in enumerate(names):
            print(f"    Wrote {st.wrote[i]} {nm} (target={st.target[i]})", file=sys.stderr)
        self.last_stats = st
        return st
