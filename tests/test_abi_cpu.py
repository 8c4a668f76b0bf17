"""No-GPU checks: the C-ABI library loads, exports every symbol include/ptq4vit_b200.h declares,
and its host-side planning validates arguments (no kernel is launched here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from ptq4vit_b200 import build, _lib
    build.build()
    return _lib.lib()


def test_header_symbols_are_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "ptq4vit_b200.h")).read()
    declared = sorted(set(re.findall(r"P4V_API[^;(]*?\b(p4v_\w+)\s*\(", hdr)))
    assert len(declared) >= 17
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    from ptq4vit_b200 import _lib
    assert sorted(_lib.EXPORTS) == declared


def _ldesc(**kw):
    from ptq4vit_b200 import _lib
    d = _lib.LinearDesc()
    base = dict(rows=6304, tokens=197, in_features=768, out_features=2304, n_V=72, n_H=24, n_a=1, w_bit=8, a_bit=8,
                eq_n=100, search_round=3, eq_alpha=0.01, eq_beta=1.2, post_gelu=0, has_bias=1, operand=0, kernel=0)
    base.update(kw)
    for k, v in base.items():
        setattr(d, k, v)
    return d


def test_linear_workspace_planning(lib):
    n = ctypes.c_size_t()
    assert lib.p4v_linear_workspace_bytes(ctypes.byref(_ldesc(operand=2)), ctypes.byref(n)) == 0
    full = n.value
    assert 5e8 < full < 5e9
    assert lib.p4v_linear_workspace_bytes(ctypes.byref(_ldesc(operand=1)), ctypes.byref(n)) == 0
    assert n.value < full          # int8 operand images are half the bf16 ones
    assert lib.p4v_linear_quant_forward_workspace_bytes(ctypes.byref(_ldesc(operand=2)), ctypes.byref(n)) == 0
    assert n.value < full / 20
    m = ctypes.c_size_t()
    assert lib.p4v_linear_score_log_floats(ctypes.byref(_ldesc()), ctypes.byref(m)) == 0
    assert m.value == 3 * (24 * 100 * 72 + 100)


@pytest.mark.parametrize("bad,msg", [
    (dict(n_H=7), "divide"), (dict(out_features=2304, n_V=100), "divide"), (dict(w_bit=9), "bit"),
    (dict(eq_n=1000), "eq_n"), (dict(rows=0), "empty"), (dict(out_features=120, n_V=5), "multiple of 16"),
    (dict(tokens=100), "tokens"),
])
def test_linear_bad_descriptors_fail_loudly(lib, bad, msg):
    n = ctypes.c_size_t()
    assert lib.p4v_linear_workspace_bytes(ctypes.byref(_ldesc(**bad)), ctypes.byref(n)) != 0
    assert msg in lib.p4v_last_error().decode()


def test_matmul_workspace_planning(lib):
    from ptq4vit_b200 import _lib
    n = ctypes.c_size_t()
    for sos, S2, S3 in ((0, 64, 197), (1, 197, 64)):
        d = _lib.MatMulDesc()
        d.batch, d.heads, d.S1, d.S2, d.S3 = 32, 12, 197, S2, S3
        d.A_bit = d.B_bit = 8; d.eq_n = 100; d.search_round = 3; d.eq_alpha = 0.01; d.eq_beta = 1.2; d.sos = sos
        assert lib.p4v_matmul_workspace_bytes(ctypes.byref(d), ctypes.byref(n)) == 0, lib.p4v_last_error()
        assert 1e8 < n.value < 1e10
    d.S2 = 0
    assert lib.p4v_matmul_workspace_bytes(ctypes.byref(d), ctypes.byref(n)) != 0


def test_null_pointers_are_rejected_before_any_launch(lib):
    d = _ldesc()
    rc = lib.p4v_linear_calibrate(ctypes.byref(d), None, None, None, None, None, None, 0, None, None, None, None)
    assert rc != 0 and "null" in lib.p4v_last_error().decode()


def test_missing_library_is_loud(monkeypatch, tmp_path):
    from ptq4vit_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU"):
        _lib.lib()


def test_cpu_tensors_are_rejected():
    import torch
    from ptq4vit_b200.quant_layers.linear import PTQSLBatchingQuantLinear
    m = PTQSLBatchingQuantLinear(32, 32, metric="hessian", eq_n=10)
    m.raw_input = torch.randn(2, 3, 32); m.raw_out = torch.randn(2, 3, 32); m.raw_grad = torch.randn(2, 3, 32)
    with pytest.raises(RuntimeError, match="CUDA"):
        m.calibration_step2()
