"""The C-ABI library loads and exports exactly what include/qtipsim.h declares (no compute)."""
import ctypes as C
import os
import re
import subprocess

import pytest

import util
from qtip_b200 import _abi
from qtip_b200.templates import CTmplPair, CTmplUnp


def header_symbols():
    with open(os.path.join(util.ROOT, "include", "qtipsim.h")) as fh:
        src = fh.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(qtipsim_[a-z_0-9]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert header_symbols() == sorted(_abi.SYMBOLS)


def test_library_exports_every_symbol():
    assert os.path.exists(_abi.LIB_PATH), "build the extension: python -c 'import __graft_entry__ as g; g.build()'"
    lib = _abi.load()
    for s in header_symbols():
        assert hasattr(lib, s), s
    assert lib.qtipsim_abi_version() == 1
    out = subprocess.run(["nm", "-D", "--defined-only", _abi.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (qtipsim_\w+)", out))
    assert exported == set(header_symbols())


def test_struct_layouts():
    """ctypes mirrors have the C layout (x86-64 SysV) of the structs in qtipsim.h."""
    assert C.sizeof(CTmplUnp) == 32 and CTmplUnp.qual.offset == 16 and CTmplUnp.opp_len.offset == 12
    assert C.sizeof(CTmplPair) == 80 and CTmplPair.fraglen.offset == 72 and CTmplPair.qual_2.offset == 48
    assert C.sizeof(_abi.RunParams) == 80 and _abi.RunParams.philox_seed.offset == 64
    assert _abi.Stats.bytes.offset == 64 and _abi.Stats.kernel_launches.offset == 136


def test_product_does_not_touch_the_oracle():
    """Nothing under qtip_b200/ or src/ may import, link or execute oracle/."""
    bad = []
    for base in ("qtip_b200", "src"):
        for dp, _, fns in os.walk(os.path.join(util.ROOT, base)):
            for fn in fns:
                if fn.endswith((".py", ".cpp", ".cu", ".h", ".cuh", ".hpp", "Makefile")):
                    with open(os.path.join(dp, fn), errors="replace") as fh:
                        txt = fh.read()
                    if re.search(r"pyoracle|liboracle|qtipsim_oracle|oracle/", txt):
                        bad.append(os.path.join(dp, fn))
    assert not bad, bad


def test_no_cpu_fallback_without_device():
    """Without a CUDA device qtipsim_create fails loudly (this container has no GPU)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from qtip_b200 import sim
    with pytest.raises(sim.QtipSimError) as ei:
        sim.Engine()
    assert ei.value.code == -2
