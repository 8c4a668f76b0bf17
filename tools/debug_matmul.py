"""Bring-up helper: matmul golden cases through one kernel variant."""
import os, sys, time, traceback
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import _cases as C  # noqa: E402

def run(name, kernel, operand):
    os.environ["P4V_KERNEL"] = kernel; os.environ["P4V_OPERAND"] = operand
    from ptq4vit_b200.quant_layers.matmul import PTQSLBatchingQuantMatMul, SoSPTQSLBatchingQuantMatMul
    sp, (A, B, Y, G), case = C.matmul_case(name)
    z, gold = C.load_golden(name)
    mod = case["mod"]
    cls = SoSPTQSLBatchingQuantMatMul if sp.sos else PTQSLBatchingQuantMatMul
    m = cls(A_bit=mod["A_bit"], B_bit=mod["B_bit"], metric="hessian", eq_alpha=sp.eq_alpha, eq_beta=sp.eq_beta, eq_n=sp.eq_n, search_round=sp.search_round)
    m.keep_scores = True
    m.raw_input, m.raw_out, m.raw_grad = [A, B], Y, G
    t0 = time.time()
    with torch.no_grad():
        m.calibration_step2()
    torch.cuda.synchronize(); dt = time.time() - t0
    errs, agree = [], []
    for gs, rs in zip(m.last_scores, gold):
        gs = gs.cpu().numpy(); gs = gs.reshape(gs.shape[0], -1); rs = np.asarray(rs).reshape(gs.shape[0], -1)
        errs.append(float(np.abs(gs - rs).max() / np.abs(rs).max())); agree.append(bool(np.array_equal(gs.argmax(0), rs.argmax(0))))
    ea = C.rel_err(torch.as_tensor(m.A_interval).cpu().numpy(), z["A_interval"]); eb = C.rel_err(m.B_interval.cpu().numpy(), z["B_interval"])
    print(f"{name:14s} {kernel:8s} {operand:5s} {dt*1e3:8.1f} ms errs {['%.1e' % e for e in errs]} agree {agree} dA {ea:.2e} dB {eb:.2e}", flush=True)

if __name__ == "__main__":
    kernel = sys.argv[1]
    for n in C.CASES["matmul"]:
        for op in ("int8", "bf16"):
            try: run(n, kernel, op)
            except Exception: print(n, kernel, op, "FAILED"); traceback.print_exc()
