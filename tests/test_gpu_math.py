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


def test_box_muller_sqrt_matches_ieee_on_its_whole_domain():
    """bm_sqrt (csrc/nig_math.cuh) == __fsqrt_rn for EVERY float of [2^-24, 2^6] and for -0: the radicand -2 log(u) of the
    Box-Muller transform cannot leave that set."""
    bad, cnt = C.c_int64(-1), C.c_int64(0)
    N.check(N.lib().nig_selftest_sqrt(0, C.byref(bad), C.byref(cnt)))
    assert cnt.value == (0x42800000 - 0x33800000 + 1) + 1, cnt.value
    assert bad.value == 0, f"{bad.value} of {cnt.value} square roots differ from IEEE sqrt"
