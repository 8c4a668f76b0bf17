"""GPU box: one ViT-B sized Linear layer through the reference (GPU) and through kernel variants; prints where the
score tables differ.  Usage: debug_fullsize.py <qkv|proj|fc1|fc2|head> [bit] [variant ...]   variant = kernel:operand[:ENV=VAL,...]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("TQDM_DISABLE", "1")
from oracle import ptq_oracle as O, ref_harness as RH  # noqa: E402

LINEAR = {"qkv": (768, 2304, 72, False, 197), "proj": (768, 768, 24, False, 197), "fc1": (768, 3072, 24, False, 197),
          "fc2": (3072, 768, 24, True, 197), "head": (768, 1000, 1, False, 0), "fc2n": (3072, 768, 24, False, 197),
          "small": (256, 256, 8, False, 197), "smallg": (512, 256, 8, True, 197)}


def main():
    name = sys.argv[1]
    bit = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 8
    variants = [a for a in sys.argv[2:] if ":" in a] or ["tcgen05:auto"]
    K, Oo, n_V, gelu, tok = LINEAR[name]
    x, W, b, y, g = O.make_linear_fixture(100 + bit + len(name), 32, tok, K, Oo, post_gelu=gelu)
    mod = dict(n_V=n_V, n_H=24, n_a=1, w_bit=bit, a_bit=bit, search_round=1)
    ref = RH.run_linear(x, W, b, y, g, post_gelu=gelu, quant_forward=False, **mod)
    rt = [s.numpy().astype(np.float64).reshape(100, -1) for s in ref["scores"]]
    print(f"reference: {ref['seconds']:.2f}s, {len(rt)} tables")
    from ptq4vit_b200.quant_layers.linear import PTQSLBatchingQuantLinear, PostGeluPTQSLBatchingQuantLinear
    for v in variants:
        parts = v.split(":")
        os.environ["P4V_KERNEL"], os.environ["P4V_OPERAND"] = parts[0], parts[1]
        envs = {}
        if len(parts) > 2:
            for kv in parts[2].split(","):
                k_, v_ = kv.split("="); envs[k_] = v_; os.environ[k_] = v_
        cls = PostGeluPTQSLBatchingQuantLinear if gelu else PTQSLBatchingQuantLinear
        m = cls(K, Oo, metric="hessian", eq_alpha=0.01, eq_beta=1.2, eq_n=100, **mod)
        m.weight.data = W.clone(); m.bias.data = b.clone(); m.cuda(); m.keep_scores = True
        m.raw_input, m.raw_out, m.raw_grad = x.cuda(), y.cuda(), g.cuda()
        with torch.no_grad():
            m.calibration_step2()
        torch.cuda.synchronize()
        gt = [s.cpu().numpy().astype(np.float64).reshape(100, -1) for s in m.last_scores]
        print(f"== {v}")
        for c in (0, 40, 80, 99):
            print("   cand", c, "ours", " ".join(f"{gt[0][c, j]:.7e}" for j in (0, 1, 18)), "| ref", " ".join(f"{rt[0][c, j]:.7e}" for j in (0, 1, 18)))
        flips = 0
        for i in (0, 1, 2, 12, 23, 24):
            if i >= len(gt):
                continue
            gsc, rsc = gt[i], rt[i]
            scale = np.abs(rsc).max()
            err = np.abs(gsc - rsc) / scale
            c, j = np.unravel_index(err.argmax(), err.shape)
            # error relative to each entry, over the better half of the candidates
            rel = np.abs(gsc - rsc) / np.abs(rsc)
            good = rsc >= np.median(rsc, axis=0, keepdims=True)
            flips_i = int((gsc.argmax(0) != rsc.argmax(0)).sum())
            print(f"  step {i:2d}: max err/tablemax {err.max():.2e} at cand {c} group {j} (got {gsc[c, j]:.6e} ref {rsc[c, j]:.6e}); "
                  f"max entrywise rel err (better half) {rel[good].max():.2e}; mean signed rel {((gsc - rsc) / np.abs(rsc)).mean():+.2e}; picks differing {flips_i}")
        for kk in envs:
            os.environ.pop(kk, None)


if __name__ == "__main__":
    main()


def truth_step0(name, bit=8, cands=(0, 40, 80, 99), groups=(0, 1, 18)):
    """fp64 evaluation of the first weight step (column block 0) for a few candidates: who is right, reference or ours?"""
    K, Oo, n_V, gelu, tok = LINEAR[name]
    x, W, b, y, g = O.make_linear_fixture(100 + bit + len(name), 32, tok, K, Oo, post_gelu=gelu)
    sp = O.LinearSpec(K, Oo, n_V=n_V, n_H=24, n_a=1, w_bit=bit, a_bit=bit, eq_n=100, search_round=1, post_gelu=gelu)
    xd, Wd, bd, yd, gd = [t.cuda() for t in (x, W, b, y, g)]
    w_int, a_int = O.linear_initial_intervals(sp, Wd, xd)
    f = O.candidate_factors(0.01, 1.2, 100).cuda()
    x_sim = O.linear_quant_input(sp, xd, a_int).double()
    out = {}
    for c in cands:
        cur = w_int.clone()
        cur[:, :, 0:1, :] = f[c] * w_int[:, :, 0:1, :]
        w_sim = O.linear_quant_weight(sp, Wd, cur).double()
        o = x_sim.reshape(-1, K) @ w_sim.t() + bd.double()
        d = (gd.double().reshape(-1, Oo) * (yd.double().reshape(-1, Oo) - o)) ** 2
        s = -d.reshape(32, -1, n_V, Oo // n_V).mean(-1).mean(1).sum(0)      # [n_V]
        out[c] = s.cpu().numpy()
    return out


if __name__ == "__main__" and os.environ.get("P4V_TRUTH"):
    name = sys.argv[1]
    t = truth_step0(name)
    for c, s in t.items():
        print("truth cand", c, " ".join(f"{s[j]:.7e}" for j in (0, 1, 18)))
