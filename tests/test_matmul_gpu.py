"""GPU parity of the head-wise MatMul search (QK and split-of-softmax SV) vs golden vectors."""
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
VARIANTS = [("tcgen05", "int8"), ("tcgen05", "bf16"), ("simt", "int8"), ("simt", "bf16")]


def _run(name, kernel, operand, monkeypatch):
    from ptq4vit_b200.quant_layers.matmul import PTQSLBatchingQuantMatMul, SoSPTQSLBatchingQuantMatMul
    monkeypatch.setenv("P4V_KERNEL", kernel)
    monkeypatch.setenv("P4V_OPERAND", operand)
    sp, (A, B, Y, G), case = C.matmul_case(name)
    z, gold = C.load_golden(name)
    cls = SoSPTQSLBatchingQuantMatMul if sp.sos else PTQSLBatchingQuantMatMul
    mod = case["mod"]
    m = cls(A_bit=mod["A_bit"], B_bit=mod["B_bit"], metric="hessian", eq_alpha=sp.eq_alpha, eq_beta=sp.eq_beta,
            eq_n=sp.eq_n, search_round=sp.search_round)
    m.keep_scores = True
    m.raw_input, m.raw_out, m.raw_grad = [A.clone(), B.clone()], Y.clone(), G.clone()
    with torch.no_grad():
        m.calibration_step2()
    torch.cuda.synchronize()
    got = [s.cpu().numpy() for s in m.last_scores]
    assert len(got) == len(gold)
    diverged = False
    for i, (gs, rs) in enumerate(zip(got, gold)):
        rs2 = np.asarray(rs).reshape(gs.shape[0], -1); gs2 = gs.reshape(gs.shape[0], -1)
        if diverged:
            break
        C.assert_scores_close(gs2, rs2, 3e-4, f"{name}/{kernel}/{operand} step {i}")
        C.assert_choice_consistent(gs2.argmax(0), rs2, 1e-4, f"{name} step {i}")
        diverged = not np.array_equal(gs2.argmax(0), rs2.argmax(0))
    a_int = torch.as_tensor(m.A_interval).cpu().numpy(); b_int = m.B_interval.cpu().numpy()
    tol = 5e-2 if diverged else 1e-6
    assert C.rel_err(a_int, z["A_interval"]) < tol
    assert C.rel_err(b_int, z["B_interval"]) < tol
    if sp.sos and not diverged:
        assert float(m.split) == float(z["split"])
    m.mode = "quant_forward"
    with torch.no_grad():
        out = m(A.cuda(), B.cuda()).cpu()
    ref = O.matmul_quant_forward(sp, A, B, torch.as_tensor(a_int), torch.as_tensor(b_int),
                                 None if not sp.sos else torch.as_tensor(float(m.split)))
    assert C.rel_err(out.numpy(), ref.numpy()) < 2e-5, "quant_forward"


@pytest.mark.parametrize("kernel,operand", VARIANTS)
@pytest.mark.parametrize("name", [n for n in C.CASES["matmul"] if not n.endswith("vits")])
def test_matmul_small(name, kernel, operand, monkeypatch):
    _run(name, kernel, operand, monkeypatch)


@pytest.mark.parametrize("name", ["mm_qk_vits", "mm_sv_vits"])
def test_matmul_vits_shapes(name, monkeypatch):
    _run(name, "tcgen05", "auto", monkeypatch)
