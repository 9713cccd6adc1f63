#!/usr/bin/env python
"""Config 5 support (BASELINE.json configs[4]): export the REFERENCE's U-Net (model/model.py, untouched) with
seeded random weights to TorchScript fp16, the way model/export_pt.py does (the real weights are a Git-LFS pointer,
SURVEY.md §0.9).  Runs only where /root/reference exists; the output is a compiled artefact of the reference and goes
to the git-ignored oracle/_ref/ (it travels to the GPU box).

    python tools/export_unet.py 1920x1080 [3840x2160 ...]

The trace bakes the TF.resize target sizes (model.py:63-64), so one file per resolution, as in the reference."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_MODEL_DIR = "/root/reference/model"


def main():
    sys.path.insert(0, REF_MODEL_DIR)
    from model import UNet  # the reference's own definition
    out_dir = os.path.join(ROOT, "oracle", "_ref")
    os.makedirs(out_dir, exist_ok=True)
    for res in sys.argv[1:] or ["1920x1080"]:
        W, H = (int(v) for v in res.split("x"))
        torch.manual_seed(0)
        model = UNet(in_channels=5, out_channels=3, features=[64, 128, 256, 512]).eval()
        with torch.no_grad():
            # heights/widths divisible by 16 never take the resize branch: the graph is size-agnostic, trace small
            th, tw = (H, W) if (H % 16 or W % 16) else (64, 64)
            traced = torch.jit.trace(model, torch.randn(1, 5, th, tw))
        traced = traced.half()
        path = os.path.join(out_dir, f"unet_{W}x{H}.pt")
        traced.save(path)
        print(path, os.path.getsize(path) >> 20, "MiB")


if __name__ == "__main__":
    main()
