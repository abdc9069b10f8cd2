"""The source-level drop-in's host pieces without a GPU: src/ds.h's reservoir sampler and
src/rnglib_shim.cpp's process-global generator, compiled into a small library by this test (tests/shim_harness.cpp),
against the unmodified reference's (oracle/_ref/libref.so): the slots add_part1 returns, the draws it consumes, i4_uni /
r4_uni_01 values, ignbin, and the generator state after each.  (tests/test_gpu_dropin.py compares the whole binary.)"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import cases
import util
from oracle import pyoracle as po

pytestmark = pytest.mark.skipif(not po.have_ref(), reason="oracle/_ref not built (needs /root/reference)")


@pytest.fixture(scope="module")
def shim(workdir):
    so = os.path.join(workdir, "libshimharness.so")
    src = os.path.join(util.ROOT, "src")
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-w", "-I" + src, "-o", so,
                    os.path.join(util.ROOT, "tests", "shim_harness.cpp"), os.path.join(src, "rnglib_shim.cpp")], check=True)
    lib = C.CDLL(so)
    lib.shim_r4_uni_01.restype = C.c_float
    lib.shim_ignbin.argtypes = [C.c_int, C.c_float]
    lib.shim_reservoir_offsets.argtypes = [C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64)]
    lib.shim_reservoir_offsets.restype = C.c_uint64
    lib.shim_reservoir_add.argtypes = [C.c_uint64, C.c_uint64, C.POINTER(C.c_int)]
    lib.shim_reservoir_add.restype = C.c_uint64
    return lib


def states(shim, ref):
    a, b, c, d = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    shim.shim_get_state(C.byref(a), C.byref(b))
    ref.ref_get_state(C.byref(c), C.byref(d))
    return (a.value, b.value), (c.value, d.value)


def test_generator(shim):
    ref = po.ref()
    shim.shim_initialize()
    ref.ref_initialize()
    s, r = states(shim, ref)
    assert s == r  # the default seeds of initialize(), rnglib.cpp:941-1009
    for seed in (1, 7, 4242):
        cg = cases.cli_seed(seed)
        shim.shim_set_seed(*cg)
        ref.ref_set_seed(*cg)
        for k in range(2000):
            if k % 3:
                assert shim.shim_i4_uni() == ref.ref_i4_uni()
            else:
                assert shim.shim_r4_uni_01() == ref.ref_r4_uni_01()
        for n, p in ((10, 0.3), (100000, 4.2e-5), (5000000, 0.0131), (30000, 0.999), (7, 0.5)):
            p = float(np.float32(p))
            for _ in range(20):
                assert shim.shim_ignbin(n, p) == ref.ref_ignbin(n, p), (n, p)
        s, r = states(shim, ref)
        assert s == r


@pytest.mark.parametrize("k,n,seed", [(1, 50, 3), (5, 5, 4), (5, 6, 5), (30, 2000, 6), (200, 150, 7), (64, 5000, 8), (0, 40, 9)])
def test_reservoir_add_part1(shim, k, n, seed):
    ref = po.ref()
    ref.ref_reservoir_offsets.argtypes = [C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64)]
    ref.ref_reservoir_offsets.restype = C.c_uint64
    cg = cases.cli_seed(seed)
    shim.shim_set_seed(*cg)
    ref.ref_set_seed(*cg)
    a, b = (C.c_uint64 * n)(), (C.c_uint64 * n)()
    assert shim.shim_reservoir_offsets(k, n, a) == ref.ref_reservoir_offsets(k, n, b) == min(k, n)
    assert list(a) == list(b)
    s, r = states(shim, ref)
    assert s == r
    # add(): item i replaces slot j = min((size_t)(r4 * n_), n_ - 1) if j < k (ds.h:421-431), restated on the same draws
    shim.shim_set_seed(*cg)
    kept = (C.c_int * max(k, 1))()
    size = shim.shim_reservoir_add(k, n, kept)
    want = []
    for i in range(n):
        if len(want) < k:
            want.append(i)
        elif b[i] < len(want):
            want[b[i]] = i
    assert list(kept)[:size] == want
    assert states(shim, ref)[0] == r
