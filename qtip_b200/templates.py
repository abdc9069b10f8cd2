"""Input-model templates: host-side mirror of the reference's template.h.

``TemplateUnpaired`` mirrors ``struct TemplateUnpaired`` (reference src/template.h:28-105),
``TemplatePaired`` mirrors ``struct TemplatePaired`` (src/template.h:123-231).  The ctypes
structures are the exact layout of ``qtipsim_tmpl_unp`` / ``qtipsim_tmpl_pair`` in
include/qtipsim.h (and of the oracle's and reference wrapper's twins, so one marshalled
array can be handed to all three).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Sequence, Tuple


@dataclass
class TemplateUnpaired:
    best_score: int
    len: int
    fw_flag: str  # 'T' / 'F'
    mate_flag: str  # '0' / '1' / '2'
    opp_len: int
    qual: bytes
    edit_xscript: bytes

    def reflen(self) -> int:  # template.h:92-95, edit_xscript.h:15-25
        return sum(self.edit_xscript.count(c) for c in (b"S", b"=", b"X", b"D"))


@dataclass
class TemplatePaired:
    score_12: int
    score_1: int
    len_1: int
    fw_flag_1: str
    qual_1: bytes
    edit_xscript_1: bytes
    score_2: int
    len_2: int
    fw_flag_2: str
    qual_2: bytes
    edit_xscript_2: bytes
    upstream1: bool
    fraglen: int


class CTmplUnp(C.Structure):
    _fields_ = [
        ("best_score", C.c_int32),
        ("len", C.c_int32),
        ("fw_flag", C.c_char),
        ("mate_flag", C.c_char),
        ("opp_len", C.c_int32),
        ("qual", C.c_char_p),
        ("xscript", C.c_char_p),
    ]


class CTmplPair(C.Structure):
    _fields_ = [
        ("score_12", C.c_int32),
        ("score_1", C.c_int32),
        ("len_1", C.c_int32),
        ("fw_flag_1", C.c_char),
        ("qual_1", C.c_char_p),
        ("xscript_1", C.c_char_p),
        ("score_2", C.c_int32),
        ("len_2", C.c_int32),
        ("fw_flag_2", C.c_char),
        ("qual_2", C.c_char_p),
        ("xscript_2", C.c_char_p),
        ("upstream1", C.c_int32),
        ("fraglen", C.c_uint64),
    ]


def marshal_unpaired(ts: Sequence[TemplateUnpaired]):
    """-> (ctypes array, keepalive)"""
    arr = (CTmplUnp * max(len(ts), 1))()
    for i, t in enumerate(ts):
        a = arr[i]
        a.best_score = t.best_score
        a.len = t.len
        a.fw_flag = t.fw_flag.encode()
        a.mate_flag = t.mate_flag.encode()
        a.opp_len = t.opp_len
        a.qual = t.qual
        a.xscript = t.edit_xscript
    return arr, list(ts)


def marshal_paired(ts: Sequence[TemplatePaired]):
    arr = (CTmplPair * max(len(ts), 1))()
    for i, t in enumerate(ts):
        a = arr[i]
        a.score_12 = t.score_12
        a.score_1 = t.score_1
        a.len_1 = t.len_1
        a.fw_flag_1 = t.fw_flag_1.encode()
        a.qual_1 = t.qual_1
        a.xscript_1 = t.edit_xscript_1
        a.score_2 = t.score_2
        a.len_2 = t.len_2
        a.fw_flag_2 = t.fw_flag_2.encode()
        a.qual_2 = t.qual_2
        a.xscript_2 = t.edit_xscript_2
        a.upstream1 = 1 if t.upstream1 else 0
        a.fraglen = t.fraglen
    return arr, list(ts)
