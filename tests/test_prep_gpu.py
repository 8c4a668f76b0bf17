"""GPU self-test of the operand-image quantiser's division shortcut (prep.cu rint_div): it must give the same
integers as the reference's `(x / interval).round_()` (reference quant_layers/linear.py:99-103), i.e. IEEE division
followed by round-half-even, for every input."""
import ctypes

import pytest

from ptq4vit_b200 import _lib

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", [1, 2026])
def test_rint_div_matches_ieee_division(seed):
    lib = _lib.lib()
    bad = ctypes.c_ulonglong(12345)
    _lib.check(lib.p4v_selftest_rint_div(1 << 31, seed, ctypes.byref(bad), None), "p4v_selftest_rint_div")
    assert bad.value == 0
