#!/usr/bin/env python
"""Generates tests/golden/ssim.npz by importing the reference's own utils/loss_utils.py (unmodified, from
/root/reference; pure torch) in THIS container: ssim values (float32 and float64 inputs) and the autograd gradient of
1 - ssim w.r.t. the first image, for seeded image pairs whose sizes are not multiples of the 16-pixel tile.
Run:  python tests/golden/make_ssim_golden.py"""
import importlib.util
import os

import numpy as np
import torch

REF = os.environ.get("DQO_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ssim.npz")


def main():
    spec = importlib.util.spec_from_file_location("ref_loss_utils", os.path.join(REF, "utils", "loss_utils.py"))
    lu = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(lu)
    out = {}
    for name, (H, W), seed in (("a", (37, 53), 1), ("b", (64, 80), 2), ("c", (21, 9), 3)):
        g = torch.Generator().manual_seed(seed)
        gt = torch.rand(3, H, W, generator=g)
        # a render-like image: the target, blurred a little, plus noise, with a black band (unrendered tiles)
        img = (0.7 * gt + 0.3 * gt.roll(2, dims=2) + 0.05 * torch.randn(3, H, W, generator=g)).clamp(0, 1)
        img[:, : H // 5] = 0
        x = img.clone().requires_grad_(True)
        val = lu.ssim(x, gt)                       # loss_update's call (mapper.py:841), unbatched [3,H,W]
        (1 - val).backward()
        x64 = img.double().clone().requires_grad_(True)
        val64 = lu.ssim(x64, gt.double())
        (1 - val64).backward()
        out[name + "_img"] = img.numpy()
        out[name + "_gt"] = gt.numpy()
        out[name + "_ssim"] = np.float32(val.item())
        out[name + "_grad"] = x.grad.numpy()
        out[name + "_ssim64"] = np.float64(val64.item())
        out[name + "_grad64"] = x64.grad.numpy()
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items() if "ssim" in k})


if __name__ == "__main__":
    main()
