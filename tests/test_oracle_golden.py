"""The CPU oracle must reproduce the reference's own outputs (golden vectors made
by tests/golden/make_golden.py from the unmodified reference classes)."""
import numpy as np
import pytest
import torch

from oracle import ptq_oracle as O
from tests import _cases as C

SMALL_LINEAR = [n for n in C.CASES["linear"] if n != "config1"]
SMALL_MATMUL = [n for n in C.CASES["matmul"] if not n.endswith("vits")]


@pytest.mark.parametrize("name", SMALL_LINEAR)
def test_linear_oracle_matches_reference(name):
    sp, (x, W, b, y, g), case = C.linear_case(name)
    z, gold_scores = C.load_golden(name)
    w_int, a_int, log = O.linear_calibrate(sp, W, b, x, y, g, return_scores=True)
    mine = C.flatten_linear_log(log)
    assert len(mine) == len(gold_scores)
    for i, (m, r) in enumerate(zip(mine, gold_scores)):
        C.assert_scores_close(m.numpy(), r, 2e-5, f"{name} step {i}")
    assert C.rel_err(w_int.numpy(), z["w_interval"]) < 1e-6
    assert C.rel_err(a_int.numpy(), z["a_interval"]) < 1e-6
    out = O.linear_quant_forward(sp, W, b, x, w_int, a_int).reshape(-1, sp.O)[:64]
    assert C.rel_err(out.numpy(), z["quant_out_sample"]) < 1e-5


@pytest.mark.parametrize("name", SMALL_MATMUL)
def test_matmul_oracle_matches_reference(name):
    sp, (A, B, Y, G), case = C.matmul_case(name)
    z, gold_scores = C.load_golden(name)
    A_int, B_int, split, log = O.matmul_calibrate(sp, A, B, Y, G, return_scores=True)
    mine = C.flatten_matmul_log(log)
    assert len(mine) == len(gold_scores)
    for i, (m, r) in enumerate(zip(mine, gold_scores)):
        C.assert_scores_close(m.numpy(), r, 2e-5, f"{name} step {i}")
    assert C.rel_err(torch.as_tensor(A_int).numpy(), z["A_interval"]) < 1e-6
    assert C.rel_err(B_int.numpy(), z["B_interval"]) < 1e-6
    if sp.sos:
        assert float(split) == float(z["split"])
    out = O.matmul_quant_forward(sp, A, B, A_int, B_int, split)
    assert C.rel_err(out.reshape(-1, out.shape[-1])[:64].numpy(), z["quant_out_sample"]) < 1e-5


def test_candidate_grid_matches_reference_formula():
    f = O.candidate_factors(0.01, 1.2, 100)
    assert f.shape == (101,) and f.dtype == torch.float32
    assert abs(float(f[0]) - 0.01) < 1e-9 and abs(float(f[100]) - 1.2) < 1e-6


def test_nonbatching_class_agrees_on_config1_golden():
    z, _ = C.load_golden("config1")
    assert np.array_equal(z["w_interval"], z["nonbatching_w_interval"])
    assert np.array_equal(z["a_interval"], z["nonbatching_a_interval"])


def test_conv_oracle_matches_reference():
    """oracle.conv_calibrate vs the golden made by the reference's ChannelwiseBatchingQuantConv2d (a_bit = 32) on the CPU."""
    import os
    z = np.load(os.path.join(C.GOLD, "conv_small.npz"))
    x, W, b, y, g = O.make_conv_fixture(31, 4, 3, 32, 16, 4)
    w_int, scores = O.conv_calibrate(W, b, x, y, g, stride=4)
    C.assert_scores_close(scores.numpy(), z["scores_000"], 2e-5, "conv_small")
    assert C.rel_err(w_int.numpy(), z["w_interval"]) < 1e-6


def test_calibrator_golden_is_complete():
    """tests/golden/calib_tiny_vit.npz (reference HessianQuantCalibrator on the tiny ViT): every module, both modes."""
    import os
    z = np.load(os.path.join(C.GOLD, "calib_tiny_vit.npz"))
    names = {k.split("|")[1] for k in z.files}
    assert len(names) == 14 and "patch_embed.proj" in names and "head" in names
    assert {k.split("|")[0] for k in z.files} == {"par", "seq"}
    # sequential mode: the reference's matmul1 sees zero gradients behind the quantized proj and picks candidate 0
    assert float(z["seq|blocks.0.attn.matmul1|A_interval"].max()) < 0.05 * float(z["par|blocks.0.attn.matmul1|A_interval"].min())
