"""Bring-up helper (GPU box): run golden linear cases through one kernel variant and print
per-step score errors against the golden tables.  Usage: debug_linear.py <kernel> [operand] [case...]"""
import os
import sys
import time
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import _cases as C  # noqa: E402


def run(name, kernel, operand):
    os.environ["P4V_KERNEL"] = kernel
    os.environ["P4V_OPERAND"] = operand
    from ptq4vit_b200.quant_layers.linear import PTQSLBatchingQuantLinear, PostGeluPTQSLBatchingQuantLinear
    sp, (x, W, b, y, g), case = C.linear_case(name)
    z, gold = C.load_golden(name)
    cls = PostGeluPTQSLBatchingQuantLinear if sp.post_gelu else PTQSLBatchingQuantLinear
    m = cls(sp.K, sp.O, bias=b is not None, metric="hessian", eq_alpha=sp.eq_alpha, eq_beta=sp.eq_beta, eq_n=sp.eq_n,
            search_round=sp.search_round, n_V=sp.n_V, n_H=sp.n_H, n_a=sp.n_a,
            w_bit=int(np.log2(sp.w_qmax)) + 1, a_bit=int(np.log2(sp.a_qmax)) + 1)
    m.weight.data = W.clone()
    if b is not None:
        m.bias.data = b.clone()
    m.cuda(); m.keep_scores = True
    m.raw_input, m.raw_out, m.raw_grad = x, y, g
    t0 = time.time()
    with torch.no_grad():
        m.calibration_step2()
    torch.cuda.synchronize()
    dt = time.time() - t0
    errs, agree = [], []
    for gs, rs in zip(m.last_scores, gold):
        gs = gs.cpu().numpy().reshape(sp.eq_n, -1); rs = np.asarray(rs).reshape(sp.eq_n, -1)
        errs.append(float(np.abs(gs - rs).max() / np.abs(rs).max()))
        agree.append(bool(np.array_equal(gs.argmax(0), rs.argmax(0))))
    ew = C.rel_err(m.w_interval.cpu().numpy(), z["w_interval"])
    ea = C.rel_err(m.a_interval.cpu().numpy(), z["a_interval"])
    print(f"{name:16s} {kernel:8s} {operand:5s} {dt*1e3:8.1f} ms  first-step err {errs[0]:.2e}  max err {max(errs):.2e} "
          f" choices agree {sum(agree)}/{len(agree)}  dW {ew:.2e} dX {ea:.2e}", flush=True)
    if errs[0] > 1e-3:
        gs = m.last_scores[0].cpu().numpy().reshape(sp.eq_n, -1); rs = np.asarray(gold[0]).reshape(sp.eq_n, -1)
        print("   got[:4]", gs[:4, :4].tolist()); print("   ref[:4]", rs[:4, :4].tolist())


if __name__ == "__main__":
    kernel = sys.argv[1]
    operands = [sys.argv[2]] if len(sys.argv) > 2 and sys.argv[2] in ("int8", "bf16", "auto") else ["int8", "bf16"]
    names = [a for a in sys.argv[2:] if a not in ("int8", "bf16", "auto")] or [n for n in C.CASES["linear"]]
    print(torch.cuda.get_device_name(0))
    for n in names:
        for op in operands:
            try:
                run(n, kernel, op)
            except Exception:
                print(f"{n} {kernel} {op} FAILED"); traceback.print_exc()
