"""GPU box: bisect the fc2 (twin-uniform, K=3072) mismatch: quant_forward vs fp64 on min-max step sizes."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ptq_oracle as O
from ptq4vit_b200.quant_layers.linear import PostGeluPTQSLBatchingQuantLinear, PTQSLBatchingQuantLinear

def run(K, Oo, n_V, n_H, gelu, seed=112):
    x, W, b, y, g = O.make_linear_fixture(seed, 32, 197, K, Oo, post_gelu=gelu)
    sp = O.LinearSpec(K, Oo, n_V=n_V, n_H=n_H, n_a=1, eq_n=100, search_round=1, post_gelu=gelu)
    xd, Wd, bd = x.cuda(), W.cuda(), b.cuda()
    w_int, a_int = O.linear_initial_intervals(sp, Wd, xd)
    cls = PostGeluPTQSLBatchingQuantLinear if gelu else PTQSLBatchingQuantLinear
    m = cls(K, Oo, metric="hessian", eq_alpha=0.01, eq_beta=1.2, eq_n=100, n_V=n_V, n_H=n_H, n_a=1, search_round=1)
    m.weight.data = W.clone(); m.bias.data = b.clone(); m.cuda()
    m.w_interval, m.a_interval, m.calibrated = w_int, a_int, True
    with torch.no_grad():
        out = m.quant_forward(xd).double().reshape(-1, Oo)
    xs = O.linear_quant_input(sp, xd, a_int).double().reshape(-1, K)
    ws = O.linear_quant_weight(sp, Wd, w_int).double()
    ref = xs @ ws.t() + bd.double()
    err = (out - ref).abs()
    print(f"K={K} O={Oo} n_V={n_V} n_H={n_H} gelu={gelu}: max abs err {err.max().item():.3e} (|ref| max {ref.abs().max().item():.3f}), "
          f"rows with err>1e-4: {(err.max(1).values > 1e-4).sum().item()} of {err.shape[0]}, cols: {(err.max(0).values > 1e-4).sum().item()} of {Oo}")
    if err.max() > 1e-4:
        r, c = np.unravel_index(err.argmax().item(), err.shape)
        print("   worst at row", r, "col", c, "got", out[r, c].item(), "ref", ref[r, c].item())
        bad_rows = torch.nonzero(err.max(1).values > 1e-4).reshape(-1)[:20].tolist()
        bad_cols = torch.nonzero(err.max(0).values > 1e-4).reshape(-1)[:20].tolist()
        print("   bad rows", bad_rows, "bad cols", bad_cols)
        # which K slab is responsible?  recompute the row with slabs removed
        xr, wr = xs[r], ws[c]
        full = (xr * wr).reshape(n_H, -1).sum(1)
        print("   slab contributions (ref):", [f"{v:.4f}" for v in full.tolist()[:8]], "... diff", (out[r, c] - ref[r, c]).item())

for env in ("int8", "bf16"):
    os.environ["P4V_OPERAND"] = env
    print("operand", env)
    run(3072, 768, 24, 24, True)
    run(3072, 768, 24, 24, False)
    run(768, 768, 24, 24, True)
    run(1024, 256, 8, 8, True)
