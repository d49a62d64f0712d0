"""CPU: pins oracle/ssim_oracle.py to the fixtures produced by the reference's own utils/loss_utils.py
(tests/golden/make_ssim_golden.py): values and autograd gradients of 1 - ssim."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ssim_oracle as so  # noqa: E402

G = np.load(os.path.join(ROOT, "tests", "golden", "ssim.npz"))


def test_ssim_value_matches_reference_fixtures():
    for n in "abc":
        img, gt = torch.from_numpy(G[n + "_img"]), torch.from_numpy(G[n + "_gt"])
        assert float(so.ssim(img, gt)) == float(G[n + "_ssim"])                       # float32: the same torch ops
        assert abs(float(so.ssim(img.double(), gt.double())) - float(G[n + "_ssim64"])) <= 1e-15
        assert abs(so.ssim_numpy(img, gt) - float(G[n + "_ssim64"])) <= 1e-12         # no convolution primitive involved


def test_ssim_gradient_matches_reference_fixtures():
    for n in "abc":
        img, gt = torch.from_numpy(G[n + "_img"]), torch.from_numpy(G[n + "_gt"])
        loss, grad = so.ssim_loss_and_grad(img, gt.permute(1, 2, 0))
        assert abs(loss - (1 - float(G[n + "_ssim64"]))) <= 1e-15
        assert np.abs(grad.numpy() - G[n + "_grad64"]).max() <= 1e-15
        # rows above the black band still receive gradient through the window; the band itself is not special-cased
        assert np.abs(G[n + "_grad64"]).max() > 0


def test_library_window_taps_are_the_bits_torch_produces():
    """No GPU needed: the C-ABI library reports its 11 window taps; they must equal, bit for bit, what the reference's
    `gaussian(11, 1.5)` (loss_utils.py:42-49, restated in the oracle with the same torch ops) evaluates to."""
    import ctypes as C
    from dqo_map_b200 import _lib
    w = (C.c_float * 11)()
    _lib.lib().dqo_ssim_window(w)
    assert np.array_equal(np.array(list(w), dtype=np.float32), so.window_1d().numpy())
