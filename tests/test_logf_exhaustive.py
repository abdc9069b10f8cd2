"""logf_glibc (qtip_b200/csrc/lecuyer.h) against the host libm's logf over EVERY positive normal float.

ignbin (ranlib.cpp:1094-1389) calls logf on the set-up and on the rare acceptance branch; the device cannot call
the host libm, so lecuyer.h restates glibc's algorithm, and replay-mode parity with the reference binary needs the two
to agree bit for bit on every argument.  2 130 706 432 floats, about 15 s on one core (the sweep is in C:
qtip_b200/csrc/hosttest.cpp::qthost_logf_sweep returns the number of mismatches)."""
import util


def test_logf_glibc_equals_libm_on_every_positive_normal_float():
    L = util.hosttest()
    assert L.qthost_logf_sweep(0x00800000, 0x7f800000, 1) == 0
