"""ignbin over continuous (n, p) (no GPU): the unmodified reference (ranlib.cpp:1094-1389), the oracle's restatement and the
product's two samplers (lecuyer.h: binom_setup + binom_sample, and the table sampler the device resolver uses) -- value of
every variate and generator state afterwards.  The fixed sweeps in test_oracle_pin.py / test_host_logic.py draw p from
nine values; here p is continuous on both sides of 0.5, n runs to 2^31 - 1, and a quarter of the points sit within 10 %
of the switch between the inversion and the BTPE branch (min(p, 1 - p) * n = 30).

p stays above 1e-6: below, the reference's inversion branch works with q = 1 - p rounded to float and qn = pow(q, n) far
from (1 - p)^n -- for n = 667 306 493, p = 3.3e-8 the probabilities it walks sum to ~1e-8 and the restart at ix > 110
(ranlib.cpp:1209) makes the call run for hours.  The restatements reproduce that faithfully; a window probability that
small needs a genome of 10^12 bases."""
import ctypes as C

import numpy as np
import pytest

import util
from oracle import pyoracle as po


@pytest.mark.skipif(not po.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("seed", [1, 2])
def test_ignbin_continuous_parameters(seed):
    o, r, L = po.orc(), po.ref(), util.hosttest()
    rng = np.random.default_rng(seed)
    branches = [0, 0]
    for _ in range(3000):
        kind = rng.integers(0, 4)
        if kind == 0:  # around the inversion / BTPE switch
            p = float(np.float32(10 ** rng.uniform(-6, -0.31)))
            n = max(1, int(30.0 / p * rng.uniform(0.9, 1.1)))
        elif kind == 1:
            p = float(np.float32(rng.uniform(0, 1)))
            n = int(10 ** rng.uniform(0, 9.3))
        elif kind == 2:
            p = float(np.float32(10 ** rng.uniform(-6, 0)))
            n = int(10 ** rng.uniform(0, 9.3))
        else:
            p = float(np.float32(1.0 - 10 ** rng.uniform(-6, -0.3)))
            n = int(10 ** rng.uniform(0, 8))
        n = min(n, 2147483647)
        if not 0.0 < p < 1.0:
            continue
        btpe = min(p, 1.0 - p) * n >= 30
        branches[btpe] += 1
        s = (int(rng.integers(1, 2147483562)), int(rng.integers(1, 2147483398)))
        st = po.OrcRng()
        o.orc_rng_set_seed(C.byref(st), *s)
        r.ref_set_seed(*s)
        s1, s2 = C.c_uint32(s[0]), C.c_uint32(s[1])
        t1, t2 = C.c_uint32(s[0]), C.c_uint32(s[1])
        for _k in range(12):
            want = r.ref_ignbin(n, p)
            assert o.orc_ignbin(C.byref(st), n, p) == want, (n, p, s)
            assert L.qthost_ignbin(C.byref(s1), C.byref(s2), n, p) == want, (n, p, s)
            if btpe:
                assert L.qthost_ignbin_tab(C.byref(t1), C.byref(t2), n, p) == want, (n, p, s)
        a, b = C.c_int(), C.c_int()
        r.ref_get_state(C.byref(a), C.byref(b))
        assert (st.cg1, st.cg2) == (s1.value, s2.value) == (a.value, b.value)
        if btpe:
            assert (t1.value, t2.value) == (a.value, b.value)
    assert min(branches) > 500
