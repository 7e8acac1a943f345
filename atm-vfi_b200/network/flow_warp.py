"""Drop-in for the reference's network/flow_warp.py (flow_warp.py:50-60): backward bilinear warp with
zero padding and align_corners=True semantics, executed by the sm_100a gather kernel."""
from __future__ import annotations

import os
import sys

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
if os.path.dirname(_HERE) not in sys.path:
    sys.path.insert(0, os.path.dirname(_HERE))

from atmvfi import _lib            # noqa: E402
from atmvfi.ops import CudaOps     # noqa: E402


def flow_warp(feature, flow, mask=False, padding_mode='zeros'):
    """feature: [B,C,H,W], flow: [B,2,H,W] in pixels (x, y).  Returns the warped tensor, and with
    ``mask=True`` also the in-bounds mask of the sampling positions (flow_warp.py:42-45)."""
    b, c, h, w = feature.size()
    assert flow.size(1) == 2
    if padding_mode != 'zeros':
        raise NotImplementedError("flow_warp: only padding_mode='zeros' (the mode every reference call site uses)")
    ops = CudaOps(feature.device, _lib.FP32)
    img = feature.detach().float().contiguous()
    fl = flow.detach().float().contiguous()
    out = torch.empty_like(img)
    with torch.cuda.device(feature.device):
        ops.flow_warp_nchw(img, fl, out)
    if not mask:
        return out
    ys, xs = torch.meshgrid(torch.arange(h, device=fl.device), torch.arange(w, device=fl.device), indexing="ij")
    gx = 2 * (xs.float() + fl[:, 0]) / (w - 1) - 1
    gy = 2 * (ys.float() + fl[:, 1]) / (h - 1) - 1
    return out, (gx >= -1) & (gy >= -1) & (gx <= 1) & (gy <= 1)
