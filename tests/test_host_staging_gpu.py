"""Pipelined host-staged search (utils/quant_calib.search_from_host): the same step sizes as searching tensors that are
already on the device (the reference moves each module's captures `.cuda()` before its search, quant_calib.py:317-356)."""
import pytest
import torch

from oracle import ptq_oracle as O

pytestmark = pytest.mark.gpu


def _linear(seed, K, Oo):
    from ptq4vit_b200.quant_layers.linear import PTQSLBatchingQuantLinear
    x, W, b, y, g = O.make_linear_fixture(seed, 4, 50, K, Oo)
    m = PTQSLBatchingQuantLinear(K, Oo, metric="hessian", eq_alpha=0.01, eq_beta=1.2, eq_n=100, search_round=2, n_V=2, n_H=2, n_a=1)
    m.weight.data = W; m.bias.data = b
    return m.cuda(), {"x": x, "y": y, "g": g}


def _matmul(seed):
    from ptq4vit_b200.quant_layers.matmul import PTQSLBatchingQuantMatMul
    A, B, Y, G = O.make_matmul_fixture(seed, 2, 3, 40, 32, 40)
    m = PTQSLBatchingQuantMatMul(metric="hessian", eq_alpha=0.01, eq_beta=1.2, eq_n=100, search_round=2)
    return m, {"A": A, "B": B, "y": Y, "g": G}


def test_search_from_host_matches_device_search():
    from ptq4vit_b200.utils.quant_calib import search_from_host
    dev = torch.device("cuda")
    mods = [_linear(3, 64, 128), _matmul(5), _linear(4, 128, 128)]
    want = []
    for m, t in mods:
        d = {k: v.to(dev) for k, v in t.items()}
        if "x" in d:
            m.raw_input, m.raw_out, m.raw_grad = d["x"], d["y"], d["g"]
        else:
            m.raw_input, m.raw_out, m.raw_grad = [d["A"], d["B"]], d["y"], d["g"]
        with torch.no_grad():
            m.calibration_step2()
        outs = [m.w_interval, m.a_interval] if "x" in d else [m.A_interval, m.B_interval]
        want.append([torch.as_tensor(o).detach().float().cpu().reshape(-1).clone() for o in outs])
        m.calibrated = False
    items = [(m, {k: v.contiguous().pin_memory() for k, v in t.items()}) for m, t in mods]
    h2d, d2h = search_from_host(items, dev)
    assert h2d == sum(v.numel() * 4 for _, t in mods for v in t.values()) and d2h > 0
    for (m, t), w in zip(mods, want):
        outs = [m.w_interval, m.a_interval] if "x" in t else [m.A_interval, m.B_interval]
        for o, ww in zip(outs, w):
            assert torch.equal(torch.as_tensor(o).detach().float().cpu().reshape(-1), ww)
