"""CPU oracle for the PTQ4ViT scale-factor search (TEST INFRASTRUCTURE ONLY).

This file is a plain torch-CPU fp32 restatement of the reference's
``calibration_step2`` interval search.  It exists so that the CUDA path can be
checked on machines where /root/reference is absent (the GPU box).  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import it; the product package
(``ptq4vit_b200``) never does.

Parity pin: the reference has no tests or golden vectors of its own
(SURVEY.md section 4).  The oracle is pinned against the reference classes
themselves, executed in the dev container on seeded inputs by
``tests/golden/make_golden.py``; the resulting vectors are committed under
``tests/golden/`` and ``tests/test_oracle_golden.py`` checks this file against
them.

Every function cites the reference lines it restates (paths relative to
/root/reference).  The arithmetic deliberately follows the reference's order
(true division, round-half-even, clamp, multiply; full fp32 GEMM per
candidate; mean over features, mean over tokens, sum over images) so that
score vectors agree to fp32 round-off.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

GELU_MIN_NEG = 0.16997124254703522  # quant_layers/linear.py:574


# --------------------------------------------------------------------------
# shared helpers
# --------------------------------------------------------------------------

def fake_quant(v: torch.Tensor, delta, lo: float, hi: float) -> torch.Tensor:
    """clamp(rne(v / delta), lo, hi) * delta  -- quant_layers/linear.py:47-48,
    :154-155, :167-168; quant_layers/matmul.py:36-37."""
    return (v / delta).round_().clamp_(lo, hi).mul_(delta)


def candidate_factors(eq_alpha: float, eq_beta: float, eq_n: int) -> torch.Tensor:
    """The eq_n+1 grid points, computed in python floats (float64) and then
    narrowed to fp32 exactly like ``torch.tensor([...])`` does
    (quant_layers/linear.py:544-545, quant_layers/matmul.py:568-569).
    Only the first eq_n are ever evaluated (linear.py:466-467)."""
    return torch.tensor([eq_alpha + i * (eq_beta - eq_alpha) / eq_n for i in range(eq_n + 1)],
                        dtype=torch.float32)


def _first_argmax(scores: torch.Tensor, dim: int = 0) -> torch.Tensor:
    return scores.argmax(dim=dim)


# --------------------------------------------------------------------------
# Linear  (quant_layers/linear.py)
# --------------------------------------------------------------------------

class LinearSpec:
    """Static description of one wrapped Linear (linear.py:98-122, :7-31)."""

    def __init__(self, in_features, out_features, n_V=1, n_H=1, n_a=1, w_bit=8, a_bit=8,
                 eq_alpha=0.01, eq_beta=1.2, eq_n=100, search_round=3, post_gelu=False):
        self.K, self.O = in_features, out_features
        self.n_V, self.n_H, self.n_a = n_V, n_H, n_a
        self.crb_rows = out_features // n_V
        self.crb_cols = in_features // n_H
        self.crb_acts = in_features // n_a
        self.w_qmax = 2 ** (w_bit - 1)
        self.a_qmax = 2 ** (a_bit - 1)
        self.eq_alpha, self.eq_beta, self.eq_n = eq_alpha, eq_beta, eq_n
        self.search_round = search_round
        self.post_gelu = post_gelu
        self.a_neg_interval = GELU_MIN_NEG / self.a_qmax  # linear.py:574


def linear_initial_intervals(sp: LinearSpec, W: torch.Tensor, x: torch.Tensor):
    """Min-max initial step sizes (linear.py:380-397; PostGelu :576-599).
    Returns w_interval [n_V,1,n_H,1], a_interval [n_a,1]."""
    w_int = W.view(sp.n_V, sp.crb_rows, sp.n_H, sp.crb_cols).abs().amax([1, 3], keepdim=True) / (sp.w_qmax - 0.5)
    xv = x.reshape(-1, sp.n_a, sp.crb_acts)
    if sp.post_gelu:
        a_int = xv.amax(dim=(0, 2)) / (sp.a_qmax - 0.5)        # signed max, linear.py:597
    else:
        a_int = xv.abs().amax(dim=(0, 2)) / (sp.a_qmax - 0.5)  # linear.py:395
    return w_int.clone(), a_int.reshape(sp.n_a, 1).clone()


def linear_quant_weight(sp: LinearSpec, W, w_interval):
    """linear.py:152-162."""
    w = fake_quant(W.view(sp.n_V, sp.crb_rows, sp.n_H, sp.crb_cols), w_interval, -sp.w_qmax, sp.w_qmax - 1)
    return w.view(sp.O, sp.K)


def linear_quant_input(sp: LinearSpec, x, a_interval):
    """linear.py:164-169; twin-uniform PostGelu variant :601-607."""
    xv = x.reshape(*x.shape[:-1], sp.n_a, sp.crb_acts)
    if sp.post_gelu:
        x_pos = (xv / a_interval).round_().clamp_(0, sp.a_qmax - 1).mul_(a_interval)
        x_neg = (xv / sp.a_neg_interval).round_().clamp_(-sp.a_qmax, 0).mul_(sp.a_neg_interval)
        return (x_pos + x_neg).reshape_as(x)
    return fake_quant(xv, a_interval, -sp.a_qmax, sp.a_qmax - 1).reshape_as(x)


def linear_quant_forward(sp: LinearSpec, W, bias, x, w_interval, a_interval):
    """Post-calibration layer output (linear.py:62-67)."""
    return F.linear(linear_quant_input(sp, x, a_interval), linear_quant_weight(sp, W, w_interval), bias)


def _hessian_score_linear(raw_out, raw_grad, out_sim, n_groups):
    """-(g*(y-yhat))^2, mean over the feature block, mean over tokens, sum over
    images (linear.py:417-423, :482-487).  out_sim: [b,*,O]; returns [n_groups]."""
    b = raw_out.shape[0]
    O = raw_out.shape[-1]
    d = (raw_grad * (raw_out - out_sim)) ** 2
    d = d.reshape(b, -1, n_groups, O // n_groups)
    s = -d
    s = s.mean(dim=-1)      # features of the row block
    s = s.mean(dim=1)       # tokens
    return s.sum(dim=0)     # images


def linear_search_w(sp: LinearSpec, W, bias, x, raw_out, raw_grad, w_interval, a_interval, w_cands,
                    chunk: int = 10, h_list=None):
    """Greedy per-column-block search of the weight step sizes
    (linear.py:455-495).  w_cands: [eq_n+1,n_V,1,n_H,1].  Returns the new
    w_interval and the list (one per h) of score matrices [eq_n,n_V]."""
    tmp = w_interval.clone().unsqueeze(0)                       # 1,n_V,1,n_H,1
    x_sim = linear_quant_input(sp, x, a_interval)
    Wv = W.view(sp.n_V, sp.crb_rows, sp.n_H, sp.crb_cols).unsqueeze(0)
    all_scores = []
    for h in (range(sp.n_H) if h_list is None else h_list):   # h_list: bounded sampling for the CPU baseline
        scores = []
        for p_st in range(0, sp.eq_n, chunk):
            p_ed = min(sp.eq_n, p_st + chunk)
            p = p_ed - p_st
            cur = tmp.repeat(p, 1, 1, 1, 1)
            cur[:, :, :, h:h + 1, :] = w_cands[p_st:p_ed, :, :, h:h + 1, :]
            w_sim = (Wv / cur).round_().clamp_(-sp.w_qmax, sp.w_qmax - 1).mul_(cur).view(-1, sp.K)
            b_sim = bias.repeat(p) if bias is not None else None
            out = F.linear(x_sim, w_sim, b_sim)                 # b,*,p*O
            for i in range(p):
                scores.append(_hessian_score_linear(raw_out, raw_grad, out[..., i * sp.O:(i + 1) * sp.O], sp.n_V))
        scores = torch.stack(scores, 0)                         # eq_n,n_V
        all_scores.append(scores)
        best = _first_argmax(scores, 0).reshape(1, -1, 1, 1, 1)
        tmp[:, :, :, h:h + 1, :] = torch.gather(w_cands[:, :, :, h:h + 1, :], 0, best)
    return tmp.squeeze(0), all_scores


def linear_search_a(sp: LinearSpec, W, bias, x, raw_out, raw_grad, w_interval, a_interval, a_cands,
                    chunk: int = 10):
    """Search of the activation step sizes (linear.py:497-533; PostGelu
    :609-642).  a_cands: [n_a,1,eq_n+1]."""
    tmp = a_interval.clone().unsqueeze(-1)                      # n_a,1,1
    w_sim = linear_quant_weight(sp, W, w_interval)
    xv = x.reshape(*x.shape[:-1], sp.n_a, sp.crb_acts)
    all_scores = []
    for a in range(sp.n_a):
        scores = []
        for c in range(sp.eq_n):
            cur = tmp[:, :, 0].clone()                          # n_a,1
            cur[a, 0] = a_cands[a, 0, c]
            if sp.post_gelu:
                x_pos = (xv / cur).round_().clamp_(0, sp.a_qmax - 1) * cur
                x_neg = (xv / sp.a_neg_interval).round_().clamp_(-sp.a_qmax, 0) * sp.a_neg_interval
                x_sim = (x_pos + x_neg).reshape_as(x)
            else:
                x_sim = ((xv / cur).round_().clamp_(-sp.a_qmax, sp.a_qmax - 1) * cur).reshape_as(x)
            out = F.linear(x_sim, w_sim, bias)
            scores.append(_hessian_score_linear(raw_out, raw_grad, out, 1)[0])
        scores = torch.stack(scores, 0)                         # eq_n
        all_scores.append(scores)
        best = int(_first_argmax(scores, 0))
        tmp[a, 0, 0] = a_cands[a, 0, best]
    return tmp.squeeze(-1), all_scores


def linear_calibrate(sp: LinearSpec, W, bias, x, raw_out, raw_grad, return_scores=False):
    """calibration_step2 of PTQSLBatchingQuantLinear (linear.py:536-555) and of
    PostGeluPTQSLBatchingQuantLinear (same driver, overridden pieces)."""
    w_int, a_int = linear_initial_intervals(sp, W, x)
    f = candidate_factors(sp.eq_alpha, sp.eq_beta, sp.eq_n).to(W.device)
    w_cands = f.view(-1, 1, 1, 1, 1) * w_int.unsqueeze(0)       # linear.py:544
    a_cands = f.view(1, 1, -1) * a_int.unsqueeze(-1)            # linear.py:545
    log = []
    for _ in range(sp.search_round):
        w_int, sw = linear_search_w(sp, W, bias, x, raw_out, raw_grad, w_int, a_int, w_cands)
        a_int, sa = linear_search_a(sp, W, bias, x, raw_out, raw_grad, w_int, a_int, a_cands)
        log.append((sw, sa))
    if return_scores:
        return w_int, a_int, log
    return w_int, a_int


# --------------------------------------------------------------------------
# MatMul  (quant_layers/matmul.py) -- head-wise groups, n_V = n_H = 1
# --------------------------------------------------------------------------

class MatMulSpec:
    """matmul.py:77-107, :391-394; SoS flag :579-593.  The Batching classes
    force one group per head (matmul.py:411-417); PTQ4ViT's config keeps
    n_V = n_H = 1 for both operands (configs/PTQ4ViT.py:36-48), which is the
    only layout this oracle (and the CUDA path) covers."""

    def __init__(self, A_bit=8, B_bit=8, eq_alpha=0.01, eq_beta=1.2, eq_n=100, search_round=3, sos=False):
        self.A_qmax = 2 ** (A_bit - 1)
        self.B_qmax = 2 ** (B_bit - 1)
        self.eq_alpha, self.eq_beta, self.eq_n = eq_alpha, eq_beta, eq_n
        self.search_round = search_round
        self.sos = sos


def matmul_initial_intervals(sp: MatMulSpec, A, B):
    """Per-head min-max (matmul.py:419-440); SoS: split=0.01 (matmul.py:354-355
    via the inherited batching init of B only)."""
    H = A.shape[1]
    A_int = A.abs().amax(dim=(0, 2, 3)).view(1, H, 1, 1, 1, 1, 1) / (sp.A_qmax - 0.5)
    B_int = B.abs().amax(dim=(0, 2, 3)).view(1, H, 1, 1, 1, 1, 1) / (sp.B_qmax - 0.5)
    return A_int, B_int


def _headwise(interval):
    return interval.view(1, -1, 1, 1)


def matmul_quant_A(sp: MatMulSpec, A, A_interval, split=None):
    """matmul.py:124-130; SoS twin-uniform :595-598."""
    if sp.sos:
        q1 = sp.A_qmax - 1
        x_high = (A.clamp(split, 1) * q1).round_().clamp_(0, q1) / q1
        x_low = (A.clamp(0, split) / A_interval).round_().clamp_(0, q1) * A_interval
        return x_high + x_low
    return fake_quant(A, _headwise(A_interval), -sp.A_qmax, sp.A_qmax - 1)


def matmul_quant_B(sp: MatMulSpec, B, B_interval):
    """matmul.py:132-138."""
    return fake_quant(B, _headwise(B_interval), -sp.B_qmax, sp.B_qmax - 1)


def matmul_quant_forward(sp: MatMulSpec, A, B, A_interval, B_interval, split=None):
    """matmul.py:140-145."""
    return matmul_quant_A(sp, A, A_interval, split) @ matmul_quant_B(sp, B, B_interval)


def _hessian_score_heads(raw_out, raw_grad, out_sim):
    """matmul.py:474-480, :511-513: mean over last dim, mean over rows, sum over
    images -> [H]."""
    s = -(raw_grad * (raw_out - out_sim)) ** 2
    s = s.mean(dim=-1)      # b,H,S1
    s = s.mean(dim=2)       # b,H
    return s.sum(dim=0)     # H


def matmul_search_A(sp, A, B, raw_out, raw_grad, A_interval, B_interval, A_cands):
    """matmul.py:483-522 (n_V_A = n_H_A = 1)."""
    B_sim = matmul_quant_B(sp, B, B_interval)
    scores = []
    for c in range(sp.eq_n):
        A_sim = fake_quant(A, _headwise(A_cands[c]), -sp.A_qmax, sp.A_qmax - 1)
        scores.append(_hessian_score_heads(raw_out, raw_grad, A_sim @ B_sim))
    scores = torch.stack(scores, 0)                             # eq_n,H
    best = _first_argmax(scores, 0).view(1, 1, -1, 1, 1, 1, 1, 1)
    new = torch.gather(A_cands, 0, best).squeeze(0)
    return new, scores


def matmul_search_B(sp, A, B, raw_out, raw_grad, A_interval, B_interval, B_cands, split=None):
    """matmul.py:524-563."""
    A_sim = matmul_quant_A(sp, A, A_interval, split)
    scores = []
    for c in range(sp.eq_n):
        B_sim = fake_quant(B, _headwise(B_cands[c]), -sp.B_qmax, sp.B_qmax - 1)
        scores.append(_hessian_score_heads(raw_out, raw_grad, A_sim @ B_sim))
    scores = torch.stack(scores, 0)
    best = _first_argmax(scores, 0).view(1, 1, -1, 1, 1, 1, 1, 1)
    new = torch.gather(B_cands, 0, best).squeeze(0)
    return new, scores


def sos_search_split(sp, A, B, raw_out, raw_grad, split_cands):
    """Split-point search of the twin-uniform softmax quantizer; note that B
    is NOT quantized here (matmul.py:600-631, :606)."""
    q1 = sp.A_qmax - 1
    scores = []
    for i in range(len(split_cands)):
        s = split_cands[i]
        cur = s / q1
        A_high = (A.clamp(s, 1) * q1).round_().clamp_(0, q1) / q1
        A_low = (A.clamp(0, s) / cur).round_().clamp_(0, q1) * cur
        out = (A_high + A_low) @ B
        sc = -(raw_grad * (raw_out - out)) ** 2
        sc = sc.mean(dim=-1)            # b,H,S1
        sc = sc.mean(dim=(1, 2))        # b
        scores.append(sc.sum(dim=0))
    scores = torch.stack(scores, 0)
    best = int(_first_argmax(scores, 0))
    split = split_cands[best]
    return split, split / q1, scores


def matmul_calibrate(sp: MatMulSpec, A, B, raw_out, raw_grad, return_scores=False):
    """calibration_step2 of PTQSLBatchingQuantMatMul (matmul.py:565-576) and of
    SoSPTQSLBatchingQuantMatMul (matmul.py:633-644)."""
    A_int, B_int = matmul_initial_intervals(sp, A, B)
    f = candidate_factors(sp.eq_alpha, sp.eq_beta, sp.eq_n).view(-1, 1, 1, 1, 1, 1, 1, 1).to(A.device)
    B_cands = f * B_int.unsqueeze(0)
    log = []
    if sp.sos:
        split = torch.tensor(0.01, device=A.device)
        A_int = split / (sp.A_qmax - 1)
        split_cands = torch.tensor([2 ** (-i) for i in range(20)], dtype=torch.float32, device=A.device)  # matmul.py:636
        for _ in range(sp.search_round):
            split, A_int, s1 = sos_search_split(sp, A, B, raw_out, raw_grad, split_cands)
            B_int, s2 = matmul_search_B(sp, A, B, raw_out, raw_grad, A_int, B_int, B_cands, split)
            log.append((s1, s2))
        res = (A_int, B_int, split)
    else:
        A_cands = f * A_int.unsqueeze(0)
        for _ in range(sp.search_round):
            A_int, s1 = matmul_search_A(sp, A, B, raw_out, raw_grad, A_int, B_int, A_cands)
            B_int, s2 = matmul_search_B(sp, A, B, raw_out, raw_grad, A_int, B_int, B_cands)
            log.append((s1, s2))
        res = (A_int, B_int, None)
    if return_scores:
        return res + (log,)
    return res


# --------------------------------------------------------------------------
# seeded synthetic layer fixtures
# --------------------------------------------------------------------------

def make_linear_fixture(seed, n_img, n_tok, K, O, post_gelu=False, grad_scale=1e-3, bias=True):
    """x, W, b, y = xW^T+b, g.  fc2-style inputs are gelu(randn)."""
    gen = torch.Generator().manual_seed(seed)
    shape = (n_img, n_tok, K) if n_tok else (n_img, K)
    x = torch.randn(*shape, generator=gen)
    if post_gelu:
        x = F.gelu(x * 1.5)
    bound = 1.0 / math.sqrt(K)
    W = (torch.rand(O, K, generator=gen) * 2 - 1) * bound
    # a few outlier rows/cols so that block-wise scales differ
    W = W * (1.0 + 2.0 * torch.rand(O, 1, generator=gen)) * (1.0 + 0.5 * torch.rand(1, K, generator=gen))
    b = ((torch.rand(O, generator=gen) * 2 - 1) * bound) if bias else None
    y = F.linear(x, W, b)
    g = torch.randn(*y.shape, generator=gen) * grad_scale
    return x, W, b, y, g


def make_matmul_fixture(seed, n_img, H, S1, S2, S3, softmax_A=False, grad_scale=1e-3):
    gen = torch.Generator().manual_seed(seed)
    A = torch.randn(n_img, H, S1, S2, generator=gen)
    if softmax_A:
        A = torch.softmax(A * 4.0, dim=-1)
    else:
        A = A * (1.0 + torch.arange(H).view(1, H, 1, 1) * 0.25)
    B = torch.randn(n_img, H, S2, S3, generator=gen) * (1.0 + torch.arange(H).view(1, H, 1, 1) * 0.1)
    Y = A @ B
    G = torch.randn(*Y.shape, generator=gen) * grad_scale
    return A, B, Y, G


# --------------------------------------------------------------------------
# Conv2d  (quant_layers/conv.py) -- channel-wise weight search, activations in FP32
# --------------------------------------------------------------------------

def conv_calibrate(W, bias, x, raw_out, raw_grad, stride, padding=0, dilation=1, w_bit=8, eq_alpha=0.01, eq_beta=1.2, eq_n=100):
    """ChannelwiseBatchingQuantConv2d.calibration_step2 with a_bit >= 32 (conv.py:591-603): min-max step size per output
    channel (:487), candidates f_c * delta0 (:593), score = -sum_images mean_positions (g*(y - conv(x, fq(w))))^2 per
    channel (:545-549), argmax per channel (:554-557).  Returns w_interval [oc,1,1,1] and the score table [eq_n, oc]."""
    q = 2 ** (w_bit - 1)
    w_int = W.abs().amax([1, 2, 3], keepdim=True) / (q - 0.5)
    f = candidate_factors(eq_alpha, eq_beta, eq_n).to(W.device)
    cands = f.view(-1, 1, 1, 1, 1) * w_int.unsqueeze(0)                  # eq_n+1, oc, 1, 1, 1
    scores = []
    for c in range(eq_n):
        w_sim = (W / cands[c]).round_().clamp_(-q, q - 1).mul_(cands[c])
        out = F.conv2d(x, w_sim, bias, stride, padding, dilation, 1)
        s = -(raw_grad * (raw_out - out)) ** 2                            # b, oc, fw, fh
        scores.append(s.mean(dim=[2, 3]).sum(dim=0))                      # oc
    scores = torch.stack(scores, 0)
    best = scores.argmax(dim=0).reshape(1, -1, 1, 1, 1)
    return torch.gather(cands, 0, best).squeeze(0), scores


def make_conv_fixture(seed, n_img, ic, oc, size, k, grad_scale=1e-3, bias=True):
    gen = torch.Generator().manual_seed(seed)
    x = torch.randn(n_img, ic, size, size, generator=gen)
    bound = 1.0 / math.sqrt(ic * k * k)
    W = (torch.rand(oc, ic, k, k, generator=gen) * 2 - 1) * bound * (1.0 + 2.0 * torch.rand(oc, 1, 1, 1, generator=gen))
    b = ((torch.rand(oc, generator=gen) * 2 - 1) * bound) if bias else None
    y = F.conv2d(x, W, b, stride=k)
    g = torch.randn(*y.shape, generator=gen) * grad_scale
    return x, W, b, y, g


# --------------------------------------------------------------------------
# Integer export  (utils/integer.py)
# --------------------------------------------------------------------------

def int_plain(x, interval, qmax):
    """integer.py:15-17 / :64-67 / :27-42: clamp(rne(x / interval), -qmax, qmax-1) as int8."""
    return (x / interval).round_().clamp_(-qmax, qmax - 1).to(torch.int8)


def int_gelu_twin(x, a_interval, a_neg_interval, qmax):
    """integer.py:51-62 (uint8 arithmetic, the +128 is added to every element)."""
    pos = (x / a_interval).round_().clamp_(0, qmax - 1).to(torch.uint8) + 128
    neg = (x / a_neg_interval).round_().clamp_(-qmax + 1, 0).abs().to(torch.uint8)
    return pos + neg


def int_sos_twin(A, split, A_interval, qmax):
    """integer.py:78-87"""
    hi = (A.clamp(split, 1) * (qmax - 1)).round_().clamp_(0, qmax - 1).to(torch.uint8) + 128
    lo = (A.clamp(0, split) / A_interval).round_().clamp_(0, qmax - 1).to(torch.uint8)
    return hi + lo
