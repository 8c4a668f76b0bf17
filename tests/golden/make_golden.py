"""Generate golden vectors by running the UNMODIFIED reference classes.

Run in the dev container only (needs /root/reference):

    python tests/golden/make_golden.py

The reference's Batching classes hard-code ``.cuda()`` (quant_layers/linear.py:391,
:461-464, :544-545; quant_layers/matmul.py:428, :493-498, :568-569).  On this
CPU-only container a harness-only shim makes ``.cuda()`` the identity.  Scores
are captured by wrapping ``torch.Tensor.argmax`` (every search step calls it
exactly once on the similarity table).  Inputs are NOT stored: the fixtures are
re-generated from seeds by ``oracle.ptq_oracle.make_*_fixture`` (torch CPU RNG).
"""
import json
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get("PTQ4VIT_REFERENCE", "/root/reference")
sys.path.insert(0, REF)

torch.Tensor.cuda = lambda self, *a, **k: self          # harness-only shim
torch.nn.Module.cuda = lambda self, *a, **k: self
torch.cuda.empty_cache = lambda: None

from quant_layers.linear import (PTQSLBatchingQuantLinear, PostGeluPTQSLBatchingQuantLinear,  # noqa: E402
                                 PTQSLQuantLinear)
from quant_layers.matmul import PTQSLBatchingQuantMatMul, SoSPTQSLBatchingQuantMatMul  # noqa: E402
from oracle.ptq_oracle import make_linear_fixture, make_matmul_fixture  # noqa: E402

_captured = []
_orig_argmax = torch.Tensor.argmax
_orig_targmax = torch.argmax


def _argmax_spy(self, *a, **k):
    _captured.append(self.detach().clone())
    return _orig_argmax(self, *a, **k)


def _targmax_spy(inp, *a, **k):
    _captured.append(inp.detach().clone())
    return _orig_targmax(inp, *a, **k)


LINEAR_CASES = {
    # name: fixture kwargs, module kwargs
    "lin_small": dict(fx=dict(seed=11, n_img=8, n_tok=50, K=96, O=128),
                      mod=dict(n_V=4, n_H=4, n_a=1, w_bit=8, a_bit=8, search_round=3)),
    "lin_w6a6_na2": dict(fx=dict(seed=12, n_img=6, n_tok=40, K=96, O=192),
                         mod=dict(n_V=4, n_H=4, n_a=2, w_bit=6, a_bit=6, search_round=2)),
    "lin_postgelu": dict(fx=dict(seed=13, n_img=8, n_tok=33, K=128, O=64, post_gelu=True),
                         mod=dict(n_V=2, n_H=4, n_a=1, w_bit=8, a_bit=8, search_round=2), post_gelu=True),
    "lin_postgelu_w6": dict(fx=dict(seed=17, n_img=4, n_tok=64, K=64, O=32, post_gelu=True),
                            mod=dict(n_V=1, n_H=1, n_a=1, w_bit=6, a_bit=6, search_round=3), post_gelu=True),
    "lin_head2d": dict(fx=dict(seed=14, n_img=32, n_tok=0, K=64, O=40),
                       mod=dict(n_V=1, n_H=1, n_a=1, w_bit=8, a_bit=8, search_round=3)),
    "lin_tinygrad": dict(fx=dict(seed=15, n_img=4, n_tok=70, K=64, O=128, grad_scale=1e-10),
                         mod=dict(n_V=2, n_H=2, n_a=1, w_bit=8, a_bit=8, search_round=1)),
    "lin_nobias_qkv": dict(fx=dict(seed=16, n_img=4, n_tok=197, K=64, O=192, bias=False),
                           mod=dict(n_V=6, n_H=2, n_a=1, w_bit=8, a_bit=8, search_round=2)),
    # BASELINE.json configs[0]
    "config1": dict(fx=dict(seed=3, n_img=32, n_tok=197, K=384, O=384),
                    mod=dict(n_V=8, n_H=8, n_a=1, w_bit=8, a_bit=8, search_round=1)),
}

MATMUL_CASES = {
    "mm_qk_small": dict(fx=dict(seed=21, n_img=4, H=3, S1=50, S2=16, S3=50), mod=dict(A_bit=8, B_bit=8, search_round=2)),
    "mm_qk_w6": dict(fx=dict(seed=22, n_img=3, H=2, S1=40, S2=32, S3=40), mod=dict(A_bit=6, B_bit=6, search_round=2)),
    "mm_sv_small": dict(fx=dict(seed=23, n_img=4, H=3, S1=50, S2=50, S3=16, softmax_A=True),
                        mod=dict(A_bit=8, B_bit=8, search_round=2), sos=True),
    "mm_sv_w6": dict(fx=dict(seed=24, n_img=3, H=2, S1=40, S2=40, S3=32, softmax_A=True),
                     mod=dict(A_bit=6, B_bit=6, search_round=2), sos=True),
    "mm_qk_vits": dict(fx=dict(seed=25, n_img=8, H=6, S1=197, S2=64, S3=197), mod=dict(A_bit=8, B_bit=8, search_round=1)),
    "mm_sv_vits": dict(fx=dict(seed=26, n_img=8, H=6, S1=197, S2=197, S3=64, softmax_A=True),
                       mod=dict(A_bit=8, B_bit=8, search_round=1), sos=True),
}

COMMON = dict(metric="hessian", eq_alpha=0.01, eq_beta=1.2, eq_n=100)


def run_linear(name, case):
    x, W, b, y, g = make_linear_fixture(**case["fx"])
    K, O = W.shape[1], W.shape[0]
    cls = PostGeluPTQSLBatchingQuantLinear if case.get("post_gelu") else PTQSLBatchingQuantLinear
    m = cls(K, O, bias=b is not None, **COMMON, **case["mod"])
    m.weight.data = W.clone()
    if b is not None:
        m.bias.data = b.clone()
    m.raw_input, m.raw_out, m.raw_grad = x.clone(), y.clone(), g.clone()
    _captured.clear()
    torch.Tensor.argmax = _argmax_spy
    t0 = time.time()
    with torch.no_grad():
        m.calibration_step2()
        dt = time.time() - t0
        torch.Tensor.argmax = _orig_argmax
        m.mode = "quant_forward"
        out_q = m(x)
    res = {"w_interval": m.w_interval.numpy(), "a_interval": m.a_interval.numpy(),
           "quant_out_sample": out_q.reshape(-1, O)[:64].numpy().copy(),
           "seconds": np.float64(dt)}
    for i, s in enumerate(_captured):
        res[f"scores_{i:03d}"] = s.numpy().astype(np.float32)
    if name == "config1":
        # a11: the non-batching class gives the same intervals on the same data
        m2 = PTQSLQuantLinear(K, O, bias=True, **COMMON, **case["mod"])
        m2.weight.data = W.clone(); m2.bias.data = b.clone()
        m2.raw_input, m2.raw_out, m2.raw_grad = x.clone(), y.clone(), g.clone()
        with torch.no_grad():
            m2.calibration_step2(x)
        res["nonbatching_w_interval"] = m2.w_interval.numpy()
        res["nonbatching_a_interval"] = m2.a_interval.numpy()
    return res


def run_matmul(name, case):
    A, B, Y, G = make_matmul_fixture(**case["fx"])
    cls = SoSPTQSLBatchingQuantMatMul if case.get("sos") else PTQSLBatchingQuantMatMul
    m = cls(**COMMON, **case["mod"])
    m.raw_input, m.raw_out, m.raw_grad = [A.clone(), B.clone()], Y.clone(), G.clone()
    _captured.clear()
    torch.argmax = _targmax_spy
    t0 = time.time()
    with torch.no_grad():
        m.calibration_step2()
        dt = time.time() - t0
        torch.argmax = _orig_targmax
        m.mode = "quant_forward"
        out_q = m(A, B)
    res = {"A_interval": torch.as_tensor(m.A_interval).numpy(), "B_interval": m.B_interval.numpy(),
           "quant_out_sample": out_q.reshape(-1, out_q.shape[-1])[:64].numpy().copy(),
           "seconds": np.float64(dt)}
    if case.get("sos"):
        res["split"] = torch.as_tensor(m.split).numpy()
    for i, s in enumerate(_captured):
        res[f"scores_{i:03d}"] = s.numpy().astype(np.float32)
    return res


def main():
    only = set(sys.argv[1:])
    meta = {"torch": torch.__version__, "threads": torch.get_num_threads(), "common": COMMON,
            "linear": {k: {kk: vv for kk, vv in v.items()} for k, v in LINEAR_CASES.items()},
            "matmul": {k: {kk: vv for kk, vv in v.items()} for k, v in MATMUL_CASES.items()}}
    for name, case in LINEAR_CASES.items():
        if only and name not in only:
            continue
        res = run_linear(name, case)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **res)
        print(name, "done in", float(res["seconds"]), "s", flush=True)
    for name, case in MATMUL_CASES.items():
        if only and name not in only:
            continue
        res = run_matmul(name, case)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **res)
        print(name, "done in", float(res["seconds"]), "s", flush=True)
    with open(os.path.join(HERE, "cases.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
