"""Debug: clock64 timeline of CTA 0 of one sweep launch (producer / MMA / epilogue warp 4)."""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from oracle import ptq_oracle as O
from ptq4vit_b200 import _lib
from ptq4vit_b200.quant_layers.linear import PTQSLBatchingQuantLinear
kind = sys.argv[1] if len(sys.argv) > 1 else "w"
lib = _lib.lib(); lib.p4v_debug_trace.argtypes = [ctypes.c_void_p]
x, W, b, y, g = O.make_linear_fixture(1, 32, 197, 768, 2304)
m = PTQSLBatchingQuantLinear(768, 2304, metric="hessian", eq_alpha=0.01, eq_beta=1.2, eq_n=100, search_round=1, n_V=72, n_H=24, n_a=1)
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
# warm
if kind == "w":
    _lib.check(lib.p4v_linear_search_w(ctypes.byref(d), P(bb), P(y2), P(g2), P(ws), 0, 2, None, st), "w")
else:
    _lib.check(lib.p4v_linear_search_a(ctypes.byref(d), P(bb), P(y2), P(g2), P(ws), 0, 1, None, st), "a")
torch.cuda.synchronize()
tr = torch.zeros(3 * 512 * 4, dtype=torch.int64, device=dev)
lib.p4v_debug_trace(ctypes.c_void_p(tr.data_ptr()))
if kind == "w":
    _lib.check(lib.p4v_linear_search_w(ctypes.byref(d), P(bb), P(y2), P(g2), P(ws), 2, 3, None, st), "w")
else:
    _lib.check(lib.p4v_linear_search_a(ctypes.byref(d), P(bb), P(y2), P(g2), P(ws), 0, 1, None, st), "a")
torch.cuda.synchronize()
lib.p4v_debug_trace(None)
t = tr.cpu().numpy().reshape(3, 512, 4)
t0 = t[1, 0, 0]
np.save(os.path.join(ROOT, "gpurun_out", f"trace_{kind}.npy"), t)
def show(role, name, cols, rng=None):
    print(name)
    for i in (rng or (list(range(0, 12)) + list(range(60, 100)))):
        e = t[role, i]
        if e[0] == 0: break
        print(i, [int(v - t0) if k < cols else int(v) for k, v in enumerate(e)], "d:", [int(e[k + 1] - e[k]) for k in range(cols - 1)],
              "period:", int(e[0] - t[role, i - 1, 0]) if i else 0)
show(0, "producer: [t_before_wait_empty, t_after_wait, t_after_issue]", 3)
show(1, "mma: [t_start, t_after_acc_empty, t_after_full, t_end]", 4)
show(2, "epilogue w4: [t_start, t_after_acc_step(release+next full), t_after_reduce, t_after_store]", 4)
for role, nm in ((0, "prod"), (1, "mma"), (2, "epi")):
    v = t[role]; k = int((v[:, 0] != 0).sum())
    if k > 60:
        per = (v[k - 1, 0] - v[40, 0]) / (k - 41)
        print(nm, "events", k, "avg cycles per event (steady)", per)
