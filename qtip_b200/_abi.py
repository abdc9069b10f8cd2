"""ctypes binding of libqtipsim's C ABI (include/qtipsim.h).

The shared library is built in-tree by ``__graft_entry__.build()`` /
``make -C qtip_b200/csrc``.  There is no fallback: if it is missing, or no CUDA device is
usable, loading / ``qtipsim_create`` raises.
"""
from __future__ import annotations

import ctypes as C
import os

from .templates import CTmplPair, CTmplUnp

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("QTIPSIM_LIB") or os.path.join(_HERE, "libqtipsim.so")

NOUT = 7
OUT_NAMES = ("u", "b_1", "b_2", "c_1", "c_2", "d_1", "d_2")
CAT_U, CAT_B, CAT_C, CAT_D = 0, 1, 2, 3
FUNC_LINEAR, FUNC_SQRT, FUNC_CONST = 1, 2, 3
RNG_REPLAY, RNG_PHILOX = 0, 1

# every symbol include/qtipsim.h declares (tests/test_abi.py checks header <-> this list <-> .so)
SYMBOLS = (
    "qtipsim_label_correctness",
    "qtipsim_result_stream_fds",
    "qtipsim_result_interleaved",
    "qtipsim_genome_save",
    "qtipsim_load_fasta_cached",
    "qtipsim_edit_xscripts",
    "qtipsim_model_begin",
    "qtipsim_model_add",
    "qtipsim_model_commit",
    "qtipsim_abi_version",
    "qtipsim_create",
    "qtipsim_destroy",
    "qtipsim_last_error",
    "qtipsim_load_fasta",
    "qtipsim_load_fasta_bytes",
    "qtipsim_num_estimated_bases",
    "qtipsim_num_records",
    "qtipsim_genome_export",
    "qtipsim_genome_adopt",
    "qtipsim_set_templates_unpaired",
    "qtipsim_set_templates_paired",
    "qtipsim_olap",
    "qtipsim_run",
    "qtipsim_result_copy",
    "qtipsim_result_device_ptr",
    "qtipsim_result_write_fds",
    "qtipsim_run_to_host",
    "qtipsim_stream_begin",
    "qtipsim_stream_run",
)


class RunParams(C.Structure):
    _fields_ = [
        ("chunksz", C.c_uint64),
        ("fraction", C.c_float),
        ("function", C.c_int32),
        ("min_u", C.c_uint64),
        ("min_c", C.c_uint64),
        ("min_d", C.c_uint64),
        ("min_b", C.c_uint64),
        ("rng_mode", C.c_int32),
        ("cg1", C.c_int32),
        ("cg2", C.c_int32),
        ("philox_seed", C.c_uint64),
        ("shard_rank", C.c_uint32),
        ("shard_count", C.c_uint32),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("wrote", C.c_uint64 * 4),
        ("target", C.c_uint64 * 4),
        ("bytes", C.c_uint64 * NOUT),
        ("draws", C.c_uint64),
        ("algorithmic_bytes", C.c_uint64),
        ("kernel_launches", C.c_uint32),
        ("ms_pack", C.c_float),
        ("ms_plan", C.c_float),
        ("ms_simulate", C.c_float),
        ("ms_total", C.c_float),
    ]


class ModelBatch(C.Structure):
    """qtipsim_model_batch (include/qtipsim.h): alignment-level arrays indexed by a1 / a2, then n records."""
    _fields_ = [
        ("best_score", C.POINTER(C.c_int32)),
        ("len", C.POINTER(C.c_int32)),
        ("fw", C.POINTER(C.c_uint8)),
        ("quals", C.c_char_p),
        ("q_off", C.POINTER(C.c_uint64)),
        ("xscripts", C.c_char_p),
        ("x_off", C.POINTER(C.c_uint64)),
        ("n_aln", C.c_uint64),
        ("n", C.c_uint64),
        ("cat", C.POINTER(C.c_uint8)),
        ("a1", C.POINTER(C.c_uint64)),
        ("a2", C.POINTER(C.c_uint64)),
        ("mate_flag", C.POINTER(C.c_uint8)),
        ("opp_len", C.POINTER(C.c_int32)),
        ("upstream1", C.POINTER(C.c_uint8)),
        ("fraglen", C.POINTER(C.c_uint64)),
    ]


_lib = None


def load() -> C.CDLL:
    """Loads libqtipsim.so (once).  Raises if the CUDA extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build the CUDA extension first "
            "(python -c 'import __graft_entry__ as g; g.build()' or make -C qtip_b200/csrc). "
            "qtip_b200 has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    H = C.c_void_p
    lib.qtipsim_abi_version.restype = C.c_int
    lib.qtipsim_create.argtypes = [C.POINTER(H), C.c_int]
    lib.qtipsim_destroy.argtypes = [H]
    lib.qtipsim_destroy.restype = None
    lib.qtipsim_last_error.argtypes = [H]
    lib.qtipsim_last_error.restype = C.c_char_p
    lib.qtipsim_load_fasta.argtypes = [H, C.POINTER(C.c_char_p), C.c_size_t]
    lib.qtipsim_load_fasta_bytes.argtypes = [H, C.c_void_p, C.c_size_t]
    lib.qtipsim_num_estimated_bases.argtypes = [H]
    lib.qtipsim_num_estimated_bases.restype = C.c_uint64
    lib.qtipsim_num_records.argtypes = [H]
    lib.qtipsim_num_records.restype = C.c_uint64
    lib.qtipsim_genome_export.argtypes = [H, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64), C.POINTER(C.c_int)]
    lib.qtipsim_genome_adopt.argtypes = [H, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]
    lib.qtipsim_set_templates_unpaired.argtypes = [H, C.c_int, C.POINTER(CTmplUnp), C.c_size_t, C.c_uint64]
    lib.qtipsim_set_templates_paired.argtypes = [H, C.c_int, C.POINTER(CTmplPair), C.c_size_t, C.c_uint64]
    lib.qtipsim_olap.argtypes = [H]
    lib.qtipsim_olap.restype = C.c_uint64
    lib.qtipsim_run.argtypes = [H, C.POINTER(RunParams), C.POINTER(Stats)]
    lib.qtipsim_result_copy.argtypes = [H, C.c_int, C.c_void_p, C.c_uint64]
    lib.qtipsim_result_device_ptr.argtypes = [H, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
    lib.qtipsim_result_write_fds.argtypes = [H, C.POINTER(C.c_int)]
    lib.qtipsim_run_to_host.argtypes = [H, C.POINTER(RunParams), C.POINTER(Stats), C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
    lib.qtipsim_stream_begin.argtypes = [H, C.c_void_p, C.c_size_t]
    lib.qtipsim_label_correctness.argtypes = [H, C.c_char_p, C.c_uint64, C.POINTER(C.c_uint64), C.c_char_p, C.c_uint64, C.POINTER(C.c_uint64),
                                              C.POINTER(C.c_int32), C.POINTER(C.c_uint64), C.c_uint64, C.c_int32, C.POINTER(C.c_int8)]
    lib.qtipsim_result_stream_fds.argtypes = [H, C.POINTER(C.c_int)]
    lib.qtipsim_result_interleaved.argtypes = [H, C.c_int, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
    lib.qtipsim_genome_save.argtypes = [H, C.c_char_p, C.c_char_p]
    lib.qtipsim_load_fasta_cached.argtypes = [H, C.c_char_p, C.c_char_p, C.POINTER(C.c_int)]
    P64 = C.POINTER(C.c_uint64)
    P32 = C.POINTER(C.c_int32)
    lib.qtipsim_edit_xscripts.argtypes = [H, C.c_char_p, C.c_uint64, P64, C.c_char_p, C.c_uint64, P64, P64, C.c_uint64, C.c_void_p, C.c_uint64,
                                          P64, P32, P32, P64, P64, P32]
    lib.qtipsim_model_begin.argtypes = [H, P64]
    lib.qtipsim_model_add.argtypes = [H, C.POINTER(ModelBatch), P32, P32]
    lib.qtipsim_model_commit.argtypes = [H, P64, P64]
    lib.qtipsim_stream_run.argtypes = [H, C.POINTER(RunParams), C.POINTER(Stats), C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
    _lib = lib
    return lib
