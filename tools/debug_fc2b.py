"""GPU box: read back the library's twin-uniform quantised activations through one-hot weights and compare with torch."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ptq_oracle as O
from ptq4vit_b200.quant_layers.linear import PostGeluPTQSLBatchingQuantLinear

K, n_H = 3072, 24
x, W, b, y, g = O.make_linear_fixture(112, 32, 197, K, 768, post_gelu=True)
xd = x.cuda()
sp = O.LinearSpec(K, 128, n_V=1, n_H=n_H, n_a=1, post_gelu=True)
a_int = (xd.reshape(-1, K).amax() / 127.5).reshape(1, 1)
xs_ref = O.linear_quant_input(sp, xd, a_int).reshape(-1, K)
rec = torch.empty_like(xs_ref)
for blk in range(K // 128):
    Wp = torch.zeros(128, K)
    for o in range(128):
        Wp[o, blk * 128 + o] = 1.0
    m = PostGeluPTQSLBatchingQuantLinear(K, 128, bias=False, n_V=1, n_H=n_H, n_a=1)
    m.weight.data = Wp; m.cuda()
    m.w_interval = torch.full((1, 1, n_H, 1), 1.0 / 127.0, device="cuda"); m.a_interval = a_int; m.calibrated = True   # w_hat = 127 * (1/127)
    with torch.no_grad():
        out = m.quant_forward(xd).reshape(-1, 128)
    rec[:, blk * 128:(blk + 1) * 128] = out
scale = float(127.0 * np.float32(1.0 / 127.0))
diff = (rec / scale - xs_ref)
bad = diff.abs() > 1e-6
print("elements differing:", int(bad.sum()), "of", bad.numel(), " max abs diff", float(diff.abs().max()))
idx = torch.nonzero(bad)[:30]
neg = 0.16997124254703522 / 128
for r, c in idx.tolist():
    v = float(xd.reshape(-1, K)[r, c])
    print(f"  x={v:+.9e} lib={float(rec[r, c] / scale):+.9e} torch={float(xs_ref[r, c]):+.9e}  x/neg={v / neg:+.6f} x/pos={v / float(a_int):+.6f}")
