"""Stage the UNMODIFIED reference under baseline/_ref/ (TEST / BASELINE INFRASTRUCTURE ONLY).

The reference (hahnyuan/PTQ4ViT) is pure Python with no setup.py / pyproject, so "installing" it is a copy of its
importable packages.  baseline/_ref/ is git-ignored (never part of this repository's history) but travels to the
GPU box with the gpurun snapshot, where /root/reference does not exist.  It is used by
  * tests/test_reference_gpu.py  -- the reference classes run on the B200 next to the CUDA path (parity at full size),
  * bench.py                      -- `--impl reference` (CPU arm) and the `reference_gpu` comparator.
Nothing under ptq4vit_b200/ imports it.

    python oracle/stage_ref.py            # dev container: copies from /root/reference (or $PTQ4VIT_REFERENCE)
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEST = os.path.join(ROOT, "baseline", "_ref")
PACKAGES = ("quant_layers", "utils", "configs")


def stage(src=None, quiet=False):
    """Copy the reference's importable packages; returns DEST, or None when no reference tree is available."""
    src = src or os.environ.get("PTQ4VIT_REFERENCE", "/root/reference")
    if not os.path.isdir(os.path.join(src, "quant_layers")):
        return DEST if os.path.isdir(os.path.join(DEST, "quant_layers")) else None
    os.makedirs(DEST, exist_ok=True)
    for pkg in PACKAGES:
        dst = os.path.join(DEST, pkg)
        if os.path.isdir(dst):
            shutil.rmtree(dst)
        shutil.copytree(os.path.join(src, pkg), dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    with open(os.path.join(DEST, "STAGED_FROM"), "w") as f:
        f.write(src + "\n")
    if not quiet:
        print(f"staged {', '.join(PACKAGES)} from {src} into {DEST}")
    return DEST


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
