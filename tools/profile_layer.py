"""Profile driver (GPU box): run the search of ONE ViT-B-shaped layer so that ncu sees a short, representative
launch sequence.  usage: profile_layer.py <qkv|proj|fc1|fc2|qk|sv> [rounds]"""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ptq_oracle as O  # fixtures only (seeded synthetic tensors)
from ptq4vit_b200.quant_layers.linear import PTQSLBatchingQuantLinear, PostGeluPTQSLBatchingQuantLinear
from ptq4vit_b200.quant_layers.matmul import PTQSLBatchingQuantMatMul, SoSPTQSLBatchingQuantMatMul

kind = sys.argv[1]
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 1
nb = int(os.environ.get("P4V_BLOCKS", "24"))
D = 768
if kind in ("qkv", "proj", "fc1", "fc2"):
    K, Oo, nV, gelu = {"qkv": (D, 3 * D, 3 * nb, False), "proj": (D, D, nb, False), "fc1": (D, 4 * D, nb, False), "fc2": (4 * D, D, nb, True)}[kind]
    x, W, b, y, g = O.make_linear_fixture(1, 32, 197, K, Oo, post_gelu=gelu)
    cls = PostGeluPTQSLBatchingQuantLinear if gelu else PTQSLBatchingQuantLinear
    m = cls(K, Oo, metric="hessian", eq_alpha=0.01, eq_beta=1.2, eq_n=100, search_round=rounds, n_V=nV, n_H=nb, n_a=1)
    m.weight.data = W; m.bias.data = b; m.cuda()
    t = [x.cuda(), y.cuda(), g.cuda()]
    def run():
        m.raw_input, m.raw_out, m.raw_grad = t
        m.calibration_step2()
else:
    sos = kind == "sv"
    A, B, Y, G = O.make_matmul_fixture(2, 32, 12, 197, 197 if sos else 64, 64 if sos else 197, softmax_A=sos)
    m = (SoSPTQSLBatchingQuantMatMul if sos else PTQSLBatchingQuantMatMul)(metric="hessian", eq_alpha=0.01, eq_beta=1.2, eq_n=100, search_round=rounds)
    t = [A.cuda(), B.cuda(), Y.cuda(), G.cuda()]
    def run():
        m.raw_input, m.raw_out, m.raw_grad = [t[0], t[1]], t[2], t[3]
        m.calibration_step2()
with torch.no_grad():
    run(); torch.cuda.synchronize()
    t0 = time.time(); run(); torch.cuda.synchronize()
print(f"{kind} rounds={rounds} blocks={nb}: {1e3 * (time.time() - t0):.2f} ms")
if os.environ.get("P4V_PROFILE_LOG"):
    import ctypes as C
    from ptq4vit_b200._lib import lib
    L = lib(); L.p4v_profile_enable(1)
    with torch.no_grad():
        run(); torch.cuda.synchronize()
    ms, n, ops = C.c_double(), C.c_longlong(), C.c_double()
    L.p4v_profile_collect(C.byref(ms), C.byref(n), C.byref(ops))
    print(f"sweep launches {n.value}, {ms.value:.3f} ms total, {ops.value / ms.value / 1e9:.1f} TFLOP/s executed")
if os.environ.get("P4V_TORCH_PROF"):
    from torch.profiler import profile, ProfilerActivity
    with torch.no_grad(), profile(activities=[ProfilerActivity.CUDA]) as prof:
        run(); torch.cuda.synchronize()
    rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
    tot = sum(e.device_time_total for e in rows)
    for e in rows[:18]:
        print(f"{e.key[:70]:70s} {e.count:5d} {e.device_time_total:10.1f} us {100 * e.device_time_total / tot:5.1f}%  avg {e.device_time_total / e.count:8.1f}")
    print(f"total device time {tot / 1e3:.2f} ms")
