"""GPU (-m gpu): the branch-free IEEE helpers of the NCC epilogue (`sqrt_fast`, `div_fast`, csrc/device_math.cuh)
against the compiler's correctly rounded `__fsqrt_rn` / `__fdiv_rn`: every result the kernels trust must be the same
bits.  The square root is checked over all 2^32 float bit patterns, the divisions over 2^30 pseudo-random pairs; the
sampling kernel's shared-reciprocal division and its relaxed guard for the constant divisors are checked the same way."""
import ctypes as C

import pytest

pytestmark = pytest.mark.gpu


def test_fast_sqrt_and_div_equal_the_ieee_intrinsics(engine):
    from b200dense import _cabi
    out = (C.c_uint64 * 8)()
    _cabi.check(engine.lib.b200d_selftest_fast_math(1 << 30, out, engine._stream()))
    sqrt_bad, sqrt_trusted, div_bad, div_trusted, sdiv_bad, sdiv_n, relaxed_bad, relaxed_n = list(out)
    assert sqrt_bad == 0 and div_bad == 0
    # sampling kernel: shared-reciprocal division inside its guard, and (quotient - 1) under the relaxed guard
    assert sdiv_bad == 0 and sdiv_n > 0.08 * (1 << 30)
    assert relaxed_bad == 0 and relaxed_n > 0.6 * (1 << 30)
    # trusted: zero (both signs) and every positive normal number from 2^-101 up = 0x7f7fffff - 0x0d000000 + 1 + 2 patterns
    assert sqrt_trusted == 0x7f7fffff - 0x0d000000 + 1 + 2
    assert div_trusted > 0.7 * (1 << 30)            # numerators within 2^+-100 (and zero): ~78 % of random bit patterns
