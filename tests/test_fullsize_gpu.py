"""Full-size (BASELINE.json configs[1]: ViT-B/224, 32 images) checks through size-independent properties.

The comparison with the reference itself at these sizes lives in tests/test_reference_gpu.py (the unmodified classes
on the same GPU); here the full-size layers are additionally checked through properties of the search itself (reference: quant_layers/linear.py:455-533, matmul.py:483-563):
  * two independent formulations of the weight steps (slab sweep on the tensor cores vs. the normal-equation
    form) must choose the same step sizes;
  * the choice is invariant under a power-of-two scaling of the gradient (argmax of -(g*(y-yhat))^2);
  * the choice is invariant under a permutation of the images (the score is a sum over images).
A different choice is tolerated only as a near-tie: at most 0.5 % of the blocks, by one grid step."""
import numpy as np
import pytest
import torch

from oracle import ptq_oracle as O   # seeded fixtures only

pytestmark = pytest.mark.gpu
GRID_STEP = (1.2 - 0.01) / 100


def _linear(K, Oo, n_V, post_gelu, seed):
    from ptq4vit_b200.quant_layers.linear import PTQSLBatchingQuantLinear, PostGeluPTQSLBatchingQuantLinear
    x, W, b, y, g = O.make_linear_fixture(seed, 32, 197, K, Oo, post_gelu=post_gelu)
    cls = PostGeluPTQSLBatchingQuantLinear if post_gelu else PTQSLBatchingQuantLinear
    m = cls(K, Oo, metric="hessian", eq_alpha=0.01, eq_beta=1.2, eq_n=100, search_round=1, n_V=n_V, n_H=24, n_a=1)
    m.weight.data = W; m.bias.data = b
    return m.cuda(), [t.cuda() for t in (x, y, g)]


def _search(m, x, y, g):
    m.raw_input, m.raw_out, m.raw_grad = x, y, g
    with torch.no_grad():
        m.calibration_step2()
    torch.cuda.synchronize()
    return m.w_interval.detach().float().cpu().numpy().reshape(-1), m.a_interval.detach().float().cpu().numpy().reshape(-1)


def _same_choice(a, b, what):
    ne = a != b
    assert ne.mean() <= 0.005, f"{what}: {ne.sum()} of {a.size} step sizes differ"
    if ne.any():     # near ties only: the neighbouring candidate f_c * delta0, f_c <= 1.2
        assert np.all(np.abs(a[ne] - b[ne]) <= 1.05 * GRID_STEP * np.maximum(a[ne], b[ne]) / 0.01), what


def test_vitb_qkv_two_formulations_agree(monkeypatch):
    m, (x, y, g) = _linear(768, 2304, 72, False, seed=21)
    monkeypatch.setenv("P4V_GRAM", "1")
    w1, a1 = _search(m, x, y, g)
    monkeypatch.setenv("P4V_GRAM", "0")
    w0, a0 = _search(m, x, y, g)
    assert np.all(np.isfinite(w1)) and np.all(w1 > 0) and np.all(a1 > 0)
    _same_choice(w1, w0, "qkv weight step sizes, normal-equation vs slab sweep")
    _same_choice(a1, a0, "qkv activation step size")


def test_vitb_qkv_gradient_scale_invariance():
    m, (x, y, g) = _linear(768, 2304, 72, False, seed=22)
    w1, a1 = _search(m, x, y, g)
    w4, a4 = _search(m, x, y, g * 4.0)
    _same_choice(w1, w4, "qkv weight step sizes under 4x gradient")
    _same_choice(a1, a4, "qkv activation step size under 4x gradient")


def test_vitb_fc2_image_permutation_invariance():
    m, (x, y, g) = _linear(3072, 768, 24, True, seed=23)
    w1, a1 = _search(m, x, y, g)
    perm = torch.randperm(x.shape[0], generator=torch.Generator().manual_seed(1)).cuda()
    w2, a2 = _search(m, x[perm].contiguous(), y[perm].contiguous(), g[perm].contiguous())
    _same_choice(w1, w2, "fc2 weight step sizes under image permutation")
    _same_choice(a1, a2, "fc2 activation step size under image permutation")


@pytest.mark.parametrize("sos", [False, True])
def test_vitb_matmul_gradient_scale_invariance(sos):
    from ptq4vit_b200.quant_layers.matmul import PTQSLBatchingQuantMatMul, SoSPTQSLBatchingQuantMatMul
    A, B, Y, G = O.make_matmul_fixture(31 + sos, 32, 12, 197, 197 if sos else 64, 64 if sos else 197, softmax_A=sos)
    m = (SoSPTQSLBatchingQuantMatMul if sos else PTQSLBatchingQuantMatMul)(metric="hessian", eq_alpha=0.01, eq_beta=1.2,
                                                                           eq_n=100, search_round=1)
    res = []
    for s in (1.0, 0.25):
        m.raw_input, m.raw_out, m.raw_grad = [A.cuda(), B.cuda()], Y.cuda(), G.cuda() * s
        with torch.no_grad():
            m.calibration_step2()
        torch.cuda.synchronize()
        res.append((torch.as_tensor(m.A_interval).detach().float().cpu().numpy().reshape(-1),
                    torch.as_tensor(m.B_interval).detach().float().cpu().numpy().reshape(-1)))
    _same_choice(res[0][0], res[1][0], "A step sizes under 0.25x gradient")
    _same_choice(res[0][1], res[1][1], "B step sizes under 0.25x gradient")
