"""Process-global generator state, mirroring the slice of rnglib the simulator's caller uses.

The reference keeps one global L'Ecuyer-Cote generator (rnglib.cpp:308-309) that the
caller seeds (``initialize()`` qtip_parse.cpp:1589, ``set_seed`` :1673), the reservoir
sampler advances (ds.h:444) and the simulator reads *and advances*.  This module holds
that state on the host side of the Python mirror; the stream itself is consumed on the
device in replay mode (qtip_b200/csrc/lecuyer.h).
"""
from __future__ import annotations

import struct

M1 = 2147483563
M2 = 2147483399
A1 = 40014
A2 = 40692

_state = [1234567890, 123456789]


def initialize() -> None:
    """rnglib.cpp:941-1009"""
    _state[0], _state[1] = 1234567890, 123456789


def set_seed(cg1: int, cg2: int) -> None:
    """rnglib.cpp:1674-1755 (range errors exit(1) in the reference)"""
    if not (1 <= cg1 < M1):
        raise SystemExit("SET_SEED - Fatal error!  Input parameter CG1 out of bounds.")
    if not (1 <= cg2 < M2):
        raise SystemExit("SET_SEED - Fatal error!  Input parameter CG2 out of bounds.")
    _state[0], _state[1] = cg1, cg2


def seed_from_cli(seed: int) -> None:
    """What qtip-parse does with its ``seed`` option (qtip_parse.cpp:1669-1674)."""
    set_seed(seed, (seed * 77) % 2147483562 + 1)


def get_state():
    """rnglib.cpp:516-570"""
    return _state[0], _state[1]


def i4_uni() -> int:
    """rnglib.cpp:573-683"""
    _state[0] = (_state[0] * A1) % M1
    _state[1] = (_state[1] * A2) % M2
    z = _state[0] - _state[1]
    if z < 1:
        z += M1 - 1
    return z


def r4_uni_01() -> float:
    """rnglib.cpp:1444-1505: (float)i * 4.656613057E-10 rounded to float"""
    f = struct.unpack("f", struct.pack("f", float(i4_uni())))[0]
    return struct.unpack("f", struct.pack("f", f * 4.656613057e-10))[0]
