"""The evaluation loop the reference's benchmark scripts share (test_xiph.py:103-148, test_snufilm.py:105-143):
pad -> model.forward -> optional flip test-time augmentation -> unpad -> PSNR / SSIM against the true middle frame.
The model is the sm_100a engine behind ``network_base.Network`` / ``network_lite.Network``; padding, flips and metrics are
torch glue around it, exactly where the reference has them."""
from typing import Dict, Iterable, Optional, Tuple

import numpy as np
import torch

from .metrics import psnr, ssim_matlab
from .utils import InputPadder


def img2tensor(img: np.ndarray) -> torch.Tensor:
    """HxWx3 uint8 RGB -> [1,3,H,W] float in [0,1] (benchmark/utils.py:83-86)."""
    if img.shape[-1] > 3:
        img = img[:, :, :3]
    return torch.from_numpy(np.ascontiguousarray(img)).permute(2, 0, 1).unsqueeze(0) / 255.0


@torch.no_grad()
def predict_middle(model, img0: torch.Tensor, img1: torch.Tensor, divisor: int = 64, TTA: bool = False) -> torch.Tensor:
    """img0, img1: [B,3,H,W] float in [0,1] on the model's device -> predicted middle frame [B,3,H,W] (un-padded).
    ``TTA``: the flip augmentation of test_xiph.py:134-138 / test_snufilm.py:125-129 - the pair flipped along H and W is
    interpolated too and the two predictions are averaged."""
    padder = InputPadder(img0.shape, divisor)
    a, b = padder.pad(img0, img1)
    pred = model.forward(a.contiguous(), b.contiguous())["I_t"]
    if TTA:
        pf = model.forward(a.flip(2).flip(3).contiguous(), b.flip(2).flip(3).contiguous())["I_t"]
        pred = (pred + pf.flip(2).flip(3)) / 2
    return padder.unpad(pred)


def evaluate_triplets(model, triplets: Iterable[Tuple[np.ndarray, np.ndarray, np.ndarray]], divisor: int = 64, TTA: bool = False,
                      device: Optional[torch.device] = None) -> Dict[str, float]:
    """triplets: (frame0, true middle, frame1) as HxWx3 uint8 RGB arrays.  Returns mean PSNR / SSIM and the count."""
    device = device or next(model.parameters()).device
    ps, ss = [], []
    for f0, ft, f1 in triplets:
        i0, it, i1 = (img2tensor(x).to(device) for x in (f0, ft, f1))
        pred = predict_middle(model, i0, i1, divisor, TTA)
        ps.append(psnr(pred, it))
        ss.append(ssim_matlab(pred, it))
    return {"psnr": float(np.mean(ps)) if ps else float("nan"), "ssim": float(np.mean(ss)) if ss else float("nan"), "n": len(ps)}


def synthetic_triplets(n: int, H: int, W: int, seed: int = 0):
    """Moving-texture triplets (smooth field + fine detail translated by a constant velocity): stand-ins with the datasets'
    shapes for machines without the Xiph / SNU-FILM files (there is no network access in the build environment)."""
    rng = np.random.default_rng(seed)
    for k in range(n):
        lo = rng.random((H // 8 + 6, W // 8 + 6, 3)).astype(np.float32)
        t = torch.from_numpy(lo).permute(2, 0, 1)[None]
        big = torch.nn.functional.interpolate(t, size=(H + 32, W + 32), mode="bicubic", align_corners=True)[0].permute(1, 2, 0).numpy()
        big = np.clip(big + 0.1 * (rng.random(big.shape).astype(np.float32) - 0.5), 0, 1)
        dy, dx = int(rng.integers(-3, 4)), int(rng.integers(-6, 7))
        crop = lambda s: (big[16 + s * dy: 16 + s * dy + H, 16 + s * dx: 16 + s * dx + W] * 255).round().astype(np.uint8)
        yield crop(-1), crop(0), crop(1)
