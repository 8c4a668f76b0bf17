"""GPU box: one module of the synthetic ViT-B at BASELINE config 3 (n_V = n_H = 24, 3 rounds) on its REAL captured tensors:
the reference class and this package step by step; prints the first search step whose pick differs and the reference's own
score gap there.  usage: debug_module_parity.py blocks.1.attn.proj"""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("TQDM_DISABLE", "1")
from oracle import ref_harness as RH  # noqa: E402
from ptq4vit_b200.configs import PTQ4ViT as cfg  # noqa: E402
from ptq4vit_b200.utils import quant_calib as Q  # noqa: E402
from ptq4vit_b200.utils.models import get_net  # noqa: E402
from ptq4vit_b200.utils.net_wrap import wrap_modules_in_net  # noqa: E402

name = sys.argv[1]
images = torch.randn(32, 3, 224, 224, generator=torch.Generator().manual_seed(3))
importlib.reload(cfg)
cfg.ptqsl_linear_kwargs.update(n_V=24, n_H=24)
net = get_net("vit_base_patch16_224", device="cuda", seed=0)
RH.add_target_noise(net, 32, 1000)
wrapped = wrap_modules_in_net(net, cfg)
cal = Q.HessianQuantCalibrator(net, wrapped, RH.ListLoader(images), sequential=False, batch_size=4)
m = wrapped[name]
raw = cal._raw_pred_softmax()
hooks = cal._hooks_for(m)
cal._fwd_bwd(raw)
for h in hooks:
    h.remove()
Q._cat_captured(m)
x, y, g = m.raw_input.clone(), m.raw_out.clone(), m.raw_grad.clone()
print("captured", tuple(x.shape), tuple(y.shape), "max|g|", float(g.abs().max()), "min over channels of max|g|", float(g.abs().amax((0, 1)).min()))
gelu = type(m).__name__.startswith("PostGelu")
mod = dict(n_V=m.n_V, n_H=m.n_H, n_a=m.n_a, w_bit=m.w_bit, a_bit=m.a_bit, search_round=m.search_round)
ref = RH.run_linear(x.cpu(), m.weight.data.cpu(), m.bias.data.cpu(), y.cpu(), g.cpu(), post_gelu=gelu, quant_forward=False, **mod)
m.keep_scores = True
with torch.no_grad():
    m.calibration_step2()
torch.cuda.synchronize()
rt = [s.numpy().astype(np.float64).reshape(100, -1) for s in ref["scores"]]
gt = [s.cpu().numpy().astype(np.float64).reshape(100, -1) for s in m.last_scores]
dw = (m.w_interval.cpu().reshape(-1) - ref["w_interval"].reshape(-1)).abs() / ref["w_interval"].reshape(-1)
print("step sizes differing:", int((dw > 2e-6).sum()), "of", dw.numel(), "max rel", float(dw.max()))
div = np.zeros(rt[0].shape[1] if rt[0].shape[1] > 1 else 1, dtype=bool)
shown = 0
for i, (a, r) in enumerate(zip(gt, rt)):
    ba, br = a.argmax(0), r.argmax(0)
    for j in np.nonzero(ba != br)[0]:
        if r.shape[1] == 1 or not div[j]:
            col = r[:, j]
            order = np.argsort(-col)
            gap = (col[br[j]] - col[ba[j]]) / abs(col[br[j]])
            err = np.abs(a[:, j] - r[:, j]).max() / np.abs(r[:, j]).max()
            print(f"step {i} (round {i // 25}, {'W h=' + str(i % 25) if i % 25 < 24 else 'X'}) group {j}: ours picks {ba[j]}, reference {br[j]}; reference gap {gap:.3e}; "
                  f"table err {err:.2e}; reference top candidates {order[:4].tolist()} scores {[f'{col[k]:.6e}' for k in order[:4]]}")
            shown += 1
            if os.environ.get("P4V_VERBOSE"):     # history of this group's column: where does the table error come from?
                for k in range(0, i + 1):
                    if gt[k].shape[1] > j:
                        e_k = np.abs(gt[k][:, j] - rt[k][:, j]) / np.abs(rt[k][:, j]).max()
                        print(f"      step {k:3d}: table err {e_k.max():.2e} at cand {int(e_k.argmax())}, picks {int(gt[k][:, j].argmax())}/{int(rt[k][:, j].argmax())}")
                    else:
                        e_k = np.abs(gt[k][:, 0] - rt[k][:, 0]) / np.abs(rt[k][:, 0]).max()
                        print(f"      step {k:3d} (X): table err {e_k.max():.2e}, picks {int(gt[k][:, 0].argmax())}/{int(rt[k][:, 0].argmax())}")
            if r.shape[1] > 1:
                div[j] = True
    if shown > 12:
        break
