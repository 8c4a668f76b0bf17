"""GPU parity: CUDA search (through the C ABI) vs the CPU oracle and the golden vectors."""
import numpy as np
import pytest
import torch

from oracle import ptq_oracle as O
from tests import _cases as C

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _cpu_reference_scalar_division(monkeypatch):
    """The golden vectors were produced by the reference running on the CPU, where torch divides the block maxima by
    the Python scalar (qmax - 0.5) the IEEE way; on the GPU the same expression is a multiplication by the reciprocal
    (one ulp apart for a third of the step sizes, see csrc/prep.cu keys_to_delta).  The library follows the GPU by
    default (tests/test_reference_gpu.py); here it is told to follow the CPU."""
    monkeypatch.setenv("P4V_SCALAR_DIV", "ieee")

SMALL = [n for n in C.CASES["linear"] if n != "config1"]
VARIANTS = [("tcgen05", "int8"), ("tcgen05", "bf16"), ("simt", "int8"), ("simt", "bf16")]


def _make_module(sp, W, b, post_gelu, monkeypatch, kernel, operand):
    from ptq4vit_b200.quant_layers.linear import PTQSLBatchingQuantLinear, PostGeluPTQSLBatchingQuantLinear
    monkeypatch.setenv("P4V_KERNEL", kernel)
    monkeypatch.setenv("P4V_OPERAND", operand)
    cls = PostGeluPTQSLBatchingQuantLinear if post_gelu else PTQSLBatchingQuantLinear
    m = cls(sp.K, sp.O, bias=b is not None, metric="hessian", eq_alpha=sp.eq_alpha, eq_beta=sp.eq_beta, eq_n=sp.eq_n,
            search_round=sp.search_round, n_V=sp.n_V, n_H=sp.n_H, n_a=sp.n_a,
            w_bit=int(np.log2(sp.w_qmax)) + 1, a_bit=int(np.log2(sp.a_qmax)) + 1)
    m.weight.data = W.clone()
    if b is not None:
        m.bias.data = b.clone()
    m.cuda()
    m.keep_scores = True
    return m


def _run_case(name, kernel, operand, monkeypatch, score_rtol=2e-4):
    sp, (x, W, b, y, g), case = C.linear_case(name)
    z, gold_scores = C.load_golden(name)
    m = _make_module(sp, W, b, sp.post_gelu, monkeypatch, kernel, operand)
    m.raw_input, m.raw_out, m.raw_grad = x.clone(), y.clone(), g.clone()
    with torch.no_grad():
        m.calibration_step2()
    torch.cuda.synchronize()
    got_scores = [s.cpu().numpy() for s in m.last_scores]
    assert len(got_scores) == len(gold_scores)
    # walk the greedy search: as long as every choice agrees with the reference the score tables must agree
    diverged = False
    for i, (gs, rs) in enumerate(zip(got_scores, gold_scores)):
        rs2 = np.asarray(rs).reshape(sp.eq_n, -1)
        gs2 = np.asarray(gs).reshape(sp.eq_n, -1)
        if not diverged:
            C.assert_scores_close(gs2, rs2, score_rtol, f"{name}/{kernel}/{operand} step {i}")
            C.assert_choice_consistent(gs2.argmax(0), rs2, 1e-4, f"{name}/{kernel}/{operand} step {i}")
            if not np.array_equal(gs2.argmax(0), rs2.argmax(0)):
                diverged = True      # a legitimate near-tie flip: later tables depend on it
    w_int, a_int = m.w_interval.cpu().numpy(), m.a_interval.cpu().numpy()
    if not diverged:
        assert C.rel_err(w_int, z["w_interval"]) < 1e-6, "w_interval"
        assert C.rel_err(a_int, z["a_interval"]) < 1e-6, "a_interval"
    else:   # a near-tie moved one block by one grid step at most
        assert C.rel_err(w_int, z["w_interval"]) < 5e-2 and C.rel_err(a_int, z["a_interval"]) < 5e-2
    # quantized layer output with the chosen scales vs the oracle's formula on the same scales
    m.mode = "quant_forward"
    with torch.no_grad():
        out = m(x.cuda()).cpu()
    ref = O.linear_quant_forward(sp, W, b, x, torch.from_numpy(w_int), torch.from_numpy(a_int))
    assert C.rel_err(out.numpy(), ref.numpy()) < 1e-5, "quant_forward"
    if not diverged:
        assert C.rel_err(out.reshape(-1, sp.O)[:64].numpy(), z["quant_out_sample"]) < 1e-3
    return m


@pytest.mark.parametrize("kernel,operand", VARIANTS)
@pytest.mark.parametrize("name", SMALL)
def test_linear_search_matches_reference(name, kernel, operand, monkeypatch):
    _run_case(name, kernel, operand, monkeypatch)


def test_config1_scale_pair_match(monkeypatch):
    """BASELINE.json configs[0]: PTQSLQuantLinear(384,384), 32x197 tokens, n_V=n_H=8."""
    _run_case("config1", "tcgen05", "auto", monkeypatch)


def test_nonbatching_class_same_result(monkeypatch):
    """PTQSLQuantLinear.calibration_step2(x) gives the batching class's intervals."""
    from ptq4vit_b200.quant_layers.linear import PTQSLQuantLinear
    sp, (x, W, b, y, g), case = C.linear_case("lin_small")
    z, _ = C.load_golden("lin_small")
    m = PTQSLQuantLinear(sp.K, sp.O, bias=True, metric="hessian", eq_alpha=sp.eq_alpha, eq_beta=sp.eq_beta,
                         eq_n=sp.eq_n, search_round=sp.search_round, n_V=sp.n_V, n_H=sp.n_H, n_a=sp.n_a)
    m.weight.data = W.clone(); m.bias.data = b.clone(); m.cuda()
    m.raw_input, m.raw_out, m.raw_grad = x.clone(), y.clone(), g.clone()
    with torch.no_grad():
        out = m.calibration_step2(x.cuda())
    assert C.rel_err(m.w_interval.cpu().numpy(), z["w_interval"]) < 5e-2
    assert out.shape == y.shape


def test_errors_are_loud():
    from ptq4vit_b200 import _lib
    from ptq4vit_b200.quant_layers.linear import PTQSLBatchingQuantLinear
    m = PTQSLBatchingQuantLinear(30, 20, n_V=4, metric="hessian", eq_n=10)   # 20/4 = 5 rows per block: not a multiple of 16
    m.cuda()
    m.raw_input = torch.randn(2, 3, 30); m.raw_out = torch.randn(2, 3, 20); m.raw_grad = torch.randn(2, 3, 20)
    with pytest.raises(_lib.NativeError):
        m.calibration_step2()
    m2 = PTQSLBatchingQuantLinear(32, 32, metric="cosine")
    m2.cuda()
    m2.raw_input = torch.randn(2, 3, 32); m2.raw_out = torch.randn(2, 3, 32); m2.raw_grad = None
    with pytest.raises(NotImplementedError):
        m2.calibration_step2()
