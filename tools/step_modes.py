"""Debug: time ONE search step (x = activation step, w = weight step 2) of a ViT-B-shaped layer under the sweep
debug modes (0 full, 1 no operand traffic/MMA, 2 no epilogue math, 3 handshakes only) in a single process.
usage: step_modes.py <qkv|proj|fc1|fc2> <x|w> [operand]"""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from oracle import ptq_oracle as O
from ptq4vit_b200 import _lib
from ptq4vit_b200.quant_layers.linear import PTQSLBatchingQuantLinear, PostGeluPTQSLBatchingQuantLinear
kind, which = sys.argv[1], sys.argv[2]
os.environ["P4V_GRAM"] = "0"
if len(sys.argv) > 3: os.environ["P4V_OPERAND"] = sys.argv[3]
D = 768; nb = 24
K, Oo, nV, gelu = {"qkv": (D, 3 * D, 3 * nb, False), "proj": (D, D, nb, False), "fc1": (D, 4 * D, nb, False), "fc2": (4 * D, D, nb, True)}[kind]
lib = _lib.lib(); lib.p4v_debug_sweep_mode.argtypes = [ctypes.c_int]
x, W, b, y, g = O.make_linear_fixture(1, 32, 197, K, Oo, post_gelu=gelu)
cls = PostGeluPTQSLBatchingQuantLinear if gelu else PTQSLBatchingQuantLinear
m = cls(K, Oo, metric="hessian", eq_alpha=0.01, eq_beta=1.2, eq_n=100, search_round=1, n_V=nV, n_H=nb, n_a=1)
m.weight.data = W; m.bias.data = b; m.cuda()
dev = torch.device("cuda")
x2, y2, g2 = [t.reshape(-1, t.shape[-1]).contiguous().cuda() for t in (x, y, g)]
d = m._desc(x2.shape[0], 197, 1, (0.01, 1.2, 100))
n = ctypes.c_size_t(); lib.p4v_linear_workspace_bytes(ctypes.byref(d), ctypes.byref(n))
ws = torch.empty(n.value, dtype=torch.uint8, device=dev)
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
w = m.weight.detach().contiguous(); bb = m.bias.detach().contiguous()
P = _lib.ptr
_lib.check(lib.p4v_linear_begin(ctypes.byref(d), P(x2), P(w), P(bb), P(y2), P(g2), P(ws), n.value, st), "begin")
def step():
    if which == "w":
        _lib.check(lib.p4v_linear_search_w(ctypes.byref(d), P(bb), P(y2), P(g2), P(ws), 2, 3, None, st), "w")
    else:
        _lib.check(lib.p4v_linear_search_a(ctypes.byref(d), P(bb), P(y2), P(g2), P(ws), 0, 1, None, st), "a")
step(); torch.cuda.synchronize()
lib.p4v_profile_enable(1)
for mode in (0, 1, 2, 3, 0):
    lib.p4v_debug_sweep_mode(mode)
    step(); step(); torch.cuda.synchronize()
    ms, nl, ops = ctypes.c_double(), ctypes.c_longlong(), ctypes.c_double()
    lib.p4v_profile_collect(ctypes.byref(ms), ctypes.byref(nl), ctypes.byref(ops))
    print(f"{kind} {which} mode {mode}: {1e3 * ms.value / nl.value:9.1f} us per sweep launch ({nl.value} launches)")
lib.p4v_debug_sweep_mode(0)
