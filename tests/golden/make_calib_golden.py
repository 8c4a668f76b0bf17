"""Golden vectors of the whole calibrator: the UNMODIFIED reference `HessianQuantCalibrator.batching_quant_calib()`
(utils/quant_calib.py:300-378) with its own utils/net_wrap.py and configs/PTQ4ViT.py on a 2-block synthetic ViT
(ptq4vit_b200.utils.models.VisionTransformer, seed 0; 8 images of 32x32, seed 3; mini-batch 4), CPU, dev container:

    TQDM_DISABLE=1 python tests/golden/make_calib_golden.py

The KL target of the no-grad pass is perturbed by a harness hook (oracle/ref_harness.add_target_noise), otherwise
the gradients are round-off noise and nothing platform independent could be pinned.  Stored: every module's chosen
step sizes, for sequential=False and sequential=True."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
os.environ.setdefault("TQDM_DISABLE", "1")

from oracle import ref_harness as RH  # noqa: E402
from ptq4vit_b200.utils.models import VisionTransformer  # noqa: E402


def main():
    out = {}
    for sequential in (False, True):
        net = VisionTransformer(**RH.TINY_VIT).eval()
        RH.add_target_noise(net, 8, RH.TINY_VIT["num_classes"])
        res, _, _ = RH.run_reference_calibrator(net, RH.tiny_images(), batch_size=4, sequential=sequential)
        for name, d in res.items():
            for key, v in d.items():
                out[f"{'seq' if sequential else 'par'}|{name}|{key}"] = v.numpy()
    np.savez_compressed(os.path.join(HERE, "calib_tiny_vit.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
