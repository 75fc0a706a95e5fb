"""Device self-tests of the math the kernels rely on for bit-exactness (run on the B200)."""
import ctypes as C

import pytest

from neorl_industrial import _native as N

pytestmark = pytest.mark.gpu


def test_fast_division_matches_ieee_on_guarded_domain():
    """DivFast (constant divisors and the variable x/y sequence, csrc/nig_math.cuh) == __fdiv_rn wherever its guard accepts."""
    bad, acc = C.c_int64(-1), C.c_int64(0)
    N.check(N.lib().nig_selftest_division(0, 1 << 31, 0x1234ABCD, C.byref(bad), C.byref(acc)))
    assert acc.value > (1 << 31), acc.value          # mode 1 and most of mode 2 are always inside the guard
    assert bad.value == 0, f"{bad.value} of {acc.value} guarded divisions differ from IEEE division"


def test_spec_normal_device_equals_oracle_on_every_word():
    """spec_normal (csrc/nig_math.cuh: inverse-CDF table + cubic) on the device == the oracle's restatement for EVERY one of
    the 2^32 input words: two wrapping checksums over the result bit patterns, computed on both sides."""
    import os, sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import oracle as O
    sums = (C.c_uint64 * 2)()
    N.check(N.lib().nig_selftest_normal(0, 0, 1, 1 << 32, sums))
    assert (sums[0], sums[1]) == O.selftest_normal(0, 1, 1 << 32)
    N.check(N.lib().nig_selftest_normal(0, 12345, 2654435761, 1 << 20, sums))
    assert (sums[0], sums[1]) == O.selftest_normal(12345, 2654435761, 1 << 20)
