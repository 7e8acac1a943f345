"""Host-side weight packing: reference state-dict tensors -> GEMM operands of the sm_100a kernels.

Done once per weight version (load_state_dict / .to()).  K ordering of every packed matrix is
(tap, source, channel) where "source" follows the order in which the reference concatenates inputs
(torch.cat on dim=1), so concatenations never have to be materialised.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import torch

from .ops import PackedGemm, round_up

Params = Dict[str, torch.Tensor]


def _pad_cols(w: torch.Tensor) -> torch.Tensor:
    k, n = w.shape
    ldw = round_up(n, 4)
    if ldw == n:
        return w.contiguous()
    out = torch.zeros(k, ldw, dtype=w.dtype, device=w.device)
    out[:, :n] = w
    return out


def _f32(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    return None if t is None else t.detach().to(torch.float32).contiguous()


def pack_conv(P: Params, name: str, split: Optional[Sequence[int]] = None, prelu: Optional[str] = None) -> PackedGemm:
    """nn.Conv2d weight [Co, Ci, k, k] -> [k*k*Ci, Co] (row = tap*Ci + ci, tap = ky*k + kx)."""
    w = P[name + ".weight"].detach().float()
    co, ci, k, _ = w.shape
    split = list(split) if split else [ci]
    assert sum(split) == ci, (name, split, ci)
    w32 = _pad_cols(w.permute(2, 3, 1, 0).reshape(k * k * ci, co))
    return PackedGemm(name, k, split, co, False, w32, _f32(P.get(name + ".bias")), _f32(P[prelu]) if prelu else None)


def pack_convp(P: Params, name: str, split: Optional[Sequence[int]] = None) -> PackedGemm:
    """The reference's conv() helper: Sequential(Conv2d, PReLU) -> name.0 / name.1."""
    return pack_conv(P, name + ".0", split, prelu=name + ".1.weight")


def pack_deconvp(P: Params, name: str, split: Optional[Sequence[int]] = None) -> PackedGemm:
    """ConvTranspose2d(k=2, s=2) weight [Ci, Co, 2, 2] -> [Ci, 4*Co], column = (dy*2+dx)*Co + co, + PReLU."""
    w = P[name + ".0.weight"].detach().float()
    ci, co = w.shape[:2]
    split = list(split) if split else [ci]
    assert sum(split) == ci
    w32 = _pad_cols(w.permute(0, 2, 3, 1).reshape(ci, 4 * co))
    return PackedGemm(name, 1, split, co, True, w32, _f32(P[name + ".0.bias"]), _f32(P[name + ".1.weight"]))


def pack_linear(P: Params, names: Sequence[str], bias: bool = True) -> PackedGemm:
    """One or several nn.Linear layers sharing their input, stacked on the output axis (q | kv -> qkv)."""
    ws = [P[n + ".weight"].detach().float() for n in names]
    w = torch.cat(ws, 0)                        # [sum Co, Ci]
    b = None
    if bias and (names[0] + ".bias") in P:
        b = torch.cat([P[n + ".bias"].detach().float() for n in names], 0).contiguous()
    return PackedGemm("+".join(names), 1, [w.shape[1]], w.shape[0], False, _pad_cols(w.t()), b, None)


def pack_dw(P: Params, name: str):
    """depth-wise Conv2d weight [C,1,3,3] -> [9][C]."""
    w = P[name + ".weight"].detach().float()
    return w.reshape(w.shape[0], 9).t().contiguous(), _f32(P[name + ".bias"])
