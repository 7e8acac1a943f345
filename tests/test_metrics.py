"""benchmark/metrics.py against the dense definition of the reference's metric code (benchmark/pytorch_msssim.py:82-135,
psnr_ssim.py:133-135): a full 11x11x11 gaussian conv3d over the replicate-padded volume."""
import math

import torch
import torch.nn.functional as F

from benchmark.metrics import psnr, ssim_matlab


def _dense_ssim(a, b, L=1.0):
    g = torch.tensor([math.exp(-(x - 5) ** 2 / (2 * 1.5 ** 2)) for x in range(11)], dtype=a.dtype)
    g = g / g.sum()
    w = (g[:, None, None] * g[None, :, None] * g[None, None, :])[None, None]
    blur = lambda t: F.conv3d(F.pad(t.unsqueeze(1), (5,) * 6, mode="replicate"), w)
    m1, m2 = blur(a), blur(b)
    s11, s22, s12 = blur(a * a) - m1 * m1, blur(b * b) - m2 * m2, blur(a * b) - m1 * m2
    c1, c2 = (0.01 * L) ** 2, (0.03 * L) ** 2
    return (((2 * m1 * m2 + c1) * (2 * s12 + c2)) / ((m1 * m1 + m2 * m2 + c1) * (s11 + s22 + c2))).mean().item()


def test_ssim_and_psnr_match_the_dense_definition():
    g = torch.Generator().manual_seed(0)
    a = torch.rand(2, 3, 40, 56, generator=g, dtype=torch.float64)
    b = (a + 0.05 * torch.randn(2, 3, 40, 56, generator=g, dtype=torch.float64)).clamp(0, 1)
    assert abs(ssim_matlab(a, b) - _dense_ssim(a, b)) < 1e-9
    assert abs(ssim_matlab(a, a) - 1.0) < 1e-12
    assert abs(psnr(a, b) - (-10 * math.log10(((a - b) ** 2).mean().item()))) < 1e-9
    assert psnr(a, a) == float("inf")


def test_input_padder_is_centred_replicate():
    from benchmark.utils import InputPadder
    x = torch.arange(2 * 3 * 5 * 7, dtype=torch.float32).reshape(2, 3, 5, 7)
    p = InputPadder(x.shape, divisor=8)
    y = p.pad(x)
    assert y.shape == (2, 3, 8, 8) and torch.equal(p.unpad(y), x)
    assert torch.equal(y[..., 0, :], y[..., 1, :]) and torch.equal(y[..., -1, :], y[..., -2, :])      # 1 row on top, 2 below
