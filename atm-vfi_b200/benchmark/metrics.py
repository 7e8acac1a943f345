"""Accuracy metrics of the reference's evaluation scripts, written from their definitions:

* ``psnr``        : -10 log10(mean squared error) on [0,1] tensors (benchmark/psnr_ssim.py:133-135, test_snufilm.py:139).
* ``ssim_matlab`` : the "matlab" SSIM of benchmark/pytorch_msssim.py:82-135 and psnr_ssim.py calculate_ssim: an 11-tap gaussian
  (sigma 1.5) applied along channel, height AND width of the replicate-padded [B,3,H,W] volume, C1 = (0.01 L)^2, C2 = (0.03 L)^2.

The 3-D gaussian window is separable, so local moments are computed with three 1-D passes instead of one dense 11^3 conv3d
(same numbers, 1/40 of the multiply-adds).  Device-agnostic torch: these run in the harness around the model, not in it.
"""
import math

import torch
import torch.nn.functional as F


def psnr(pred: torch.Tensor, target: torch.Tensor) -> float:
    d = pred.double() - target.double()
    mse = (d * d).mean().item()
    return float("inf") if mse == 0 else -10.0 * math.log10(mse)


def _gauss(n: int, sigma: float, device, dtype) -> torch.Tensor:
    x = torch.arange(n, device=device, dtype=dtype) - n // 2
    g = torch.exp(-(x * x) / (2.0 * sigma * sigma))
    return g / g.sum()


def _blur3(v: torch.Tensor, g: torch.Tensor) -> torch.Tensor:
    """v: [B,1,D,H,W] already replicate-padded by len(g)//2 on D, H and W -> valid gaussian mean along the three axes."""
    n = g.numel()
    v = F.conv3d(v, g.view(1, 1, n, 1, 1))
    v = F.conv3d(v, g.view(1, 1, 1, n, 1))
    return F.conv3d(v, g.view(1, 1, 1, 1, n))


def ssim_matlab(img1: torch.Tensor, img2: torch.Tensor, val_range: float = 1.0, window_size: int = 11) -> float:
    """img1, img2: [B,C,H,W] in [0, val_range]."""
    assert img1.shape == img2.shape and img1.dim() == 4
    n = min(window_size, img1.shape[-2], img1.shape[-1])
    g = _gauss(n, 1.5, img1.device, img1.dtype)
    pad = (5,) * 6                                           # the reference pads by 5 regardless of the window
    vol = lambda t: F.pad(t.unsqueeze(1), pad, mode="replicate")
    a, b = img1, img2
    mu1, mu2 = _blur3(vol(a), g), _blur3(vol(b), g)
    s11 = _blur3(vol(a * a), g) - mu1 * mu1
    s22 = _blur3(vol(b * b), g) - mu2 * mu2
    s12 = _blur3(vol(a * b), g) - mu1 * mu2
    c1, c2 = (0.01 * val_range) ** 2, (0.03 * val_range) ** 2
    m = ((2 * mu1 * mu2 + c1) * (2 * s12 + c2)) / ((mu1 * mu1 + mu2 * mu2 + c1) * (s11 + s22 + c2))
    return float(m.mean().item())
