"""CPU restatement of the SSIM term of Mapping.loss_update -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's baseline legs may import this module; the product path never does.
Two forms of the same arithmetic:
  * `ssim` -- torch, any dtype (float64 on the CPU is the arbiter of the GPU tests), differentiable through autograd:
    utils/loss_utils.py:42-99 restated (1-D Gaussian of 11 taps, sigma 1.5, normalised; 2-D window = outer product; five
    depthwise zero-padded convolutions; C1 = 0.01^2, C2 = 0.03^2; mean over every element);
  * `ssim_numpy` -- plain numpy float64 loops over the 121 taps, no convolution primitive involved.
Pinned against tests/golden/ssim.npz, written by tests/golden/make_ssim_golden.py from the reference's own
utils/loss_utils.py (imported unmodified from /root/reference): values and autograd gradients."""
from math import exp

import numpy as np
import torch
import torch.nn.functional as F


def window_1d(window_size=11, sigma=1.5, dtype=torch.float32):
    """loss_utils.py:42-49: float32 tensor of the Python-float taps, divided by its sum."""
    g = torch.tensor([exp(-((x - window_size // 2) ** 2) / float(2 * sigma ** 2)) for x in range(window_size)],
                     dtype=torch.float32)
    return (g / g.sum()).to(dtype)


def ssim(img1, img2, window_size=11):
    """loss_utils.py:61-99, size_average=True.  img1, img2: [3,H,W] (the reference passes unbatched images)."""
    channel = img1.shape[-3]
    w1 = window_1d(window_size, 1.5, torch.float32).unsqueeze(1)
    w2 = w1.mm(w1.t()).float()                       # loss_utils.py:52-58: the outer product is rounded to float32
    window = w2.expand(channel, 1, window_size, window_size).contiguous().to(img1.dtype).to(img1.device)
    pad = window_size // 2
    conv = lambda t: F.conv2d(t, window, padding=pad, groups=channel)
    mu1, mu2 = conv(img1), conv(img2)
    mu1_sq, mu2_sq, mu1_mu2 = mu1.pow(2), mu2.pow(2), mu1 * mu2
    sigma1_sq = conv(img1 * img1) - mu1_sq
    sigma2_sq = conv(img2 * img2) - mu2_sq
    sigma12 = conv(img1 * img2) - mu1_mu2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    ssim_map = ((2 * mu1_mu2 + C1) * (2 * sigma12 + C2)) / ((mu1_sq + mu2_sq + C1) * (sigma1_sq + sigma2_sq + C2))
    return ssim_map.mean()


def ssim_loss_and_grad(image, gt_hwc, dtype=torch.float64):
    """(1 - ssim, d(1 - ssim)/d image) as loss_update uses it (mapper.py:839-841): image [3,H,W], gt [H,W,3]."""
    x = image.detach().to("cpu", dtype).clone().requires_grad_(True)
    y = gt_hwc.detach().to("cpu", dtype).permute(2, 0, 1)
    loss = 1 - ssim(x, y)
    loss.backward()
    return float(loss.detach()), x.grad


def ssim_numpy(img1, img2, window_size=11):
    """The same value with explicit loops over the taps (float64)."""
    a = np.asarray(img1, dtype=np.float64)
    b = np.asarray(img2, dtype=np.float64)
    C, H, W = a.shape
    w1 = window_1d(window_size).numpy().astype(np.float32)
    w2 = np.outer(w1, w1).astype(np.float32).astype(np.float64)
    r = window_size // 2

    def conv(t):
        p = np.zeros((C, H + 2 * r, W + 2 * r))
        p[:, r:r + H, r:r + W] = t
        out = np.zeros((C, H, W))
        for i in range(window_size):
            for j in range(window_size):
                out += w2[i, j] * p[:, i:i + H, j:j + W]
        return out

    mu1, mu2 = conv(a), conv(b)
    s1, s2, s12 = conv(a * a) - mu1 ** 2, conv(b * b) - mu2 ** 2, conv(a * b) - mu1 * mu2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    m = ((2 * mu1 * mu2 + C1) * (2 * s12 + C2)) / ((mu1 ** 2 + mu2 ** 2 + C1) * (s1 + s2 + C2))
    return float(m.mean())
