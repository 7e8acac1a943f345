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


# ------------------------------------------------------------------------------------------------
# tcgen05 (TF32) operand layout - must agree with atmvfi_gemm_conv_plan() in csrc/gemm_conv_tc.cu
# ------------------------------------------------------------------------------------------------
TC_CHUNK = 32          # fp32 elements per K step (one 128-byte swizzle row)
TC_MAX_N = 256


def tc_layout(split: Sequence[int], ksize: int, cout: int, shuffle: bool) -> dict:
    chunks = [-(-c // TC_CHUNK) for c in split]
    ktc = ksize * ksize * sum(chunks) * TC_CHUNK
    cq_pad = round_up(cout, 32)
    n_need = 4 * cq_pad if shuffle else round_up(cout, 16)
    n_tiles = -(-n_need // TC_MAX_N)
    block_n = round_up(-(-n_need // n_tiles), 32 if shuffle else 16)
    return dict(chunks=chunks, ktc=ktc, cq_pad=cq_pad, n_tiles=n_tiles, block_n=block_n, n_pad=n_tiles * block_n)


def round_tf32(t: torch.Tensor) -> torch.Tensor:
    """fp32 -> nearest TF32 value (10-bit mantissa), ties away from zero (cvt.rna.tf32.f32)."""
    bits = t.contiguous().view(torch.int32)
    bits = (bits + 0x1000) & ~0x1FFF
    return bits.view(torch.float32)


def pack_tc(w: PackedGemm) -> torch.Tensor:
    """[K, ldw] fp32 GEMM operand -> [N_pad][K_tc] K-major: every (tap, source) block is padded to whole
    32-channel chunks (TMA zero-fills the matching activation channels), ConvTranspose column blocks are
    padded to a multiple of 32 columns; values pre-rounded to TF32."""
    return round_tf32(_tc_matrix(w, tc_layout(w.split, w.ksize, w.Cout, w.shuffle)))


def _tc_matrix(w: PackedGemm, L: dict) -> torch.Tensor:
    """The un-rounded [N_pad][K_tc] matrix of pack_tc."""
    taps, ctot = w.ksize * w.ksize, sum(w.split)
    n_true = 4 * w.Cout if w.shuffle else w.Cout
    src = w.w32[:, :n_true].reshape(taps, ctot, n_true)
    parts, c0 = [], 0
    for c, ch in zip(w.split, L["chunks"]):
        blk = src[:, c0 : c0 + c]
        if ch * TC_CHUNK != c:
            blk = torch.nn.functional.pad(blk, (0, 0, 0, ch * TC_CHUNK - c))
        parts.append(blk)
        c0 += c
    kmat = torch.cat(parts, 1).reshape(L["ktc"], n_true)           # [K_tc, n_true]
    out = torch.zeros(L["n_pad"], L["ktc"], dtype=torch.float32, device=w.w32.device)
    if w.shuffle:
        for q in range(4):
            out[q * L["cq_pad"] : q * L["cq_pad"] + w.Cout] = kmat[:, q * w.Cout : (q + 1) * w.Cout].t()
    else:
        out[: w.Cout] = kmat.t()
    return out


def pack_tc_x3(w: PackedGemm) -> torch.Tensor:
    """3xTF32 operand: the pack_tc layout with every 32-column K chunk stored twice - hi = tf32(w) then lo = tf32(w - hi) -
    so [N_pad][2 * K_tc]; chunk kb occupies columns [64 kb, 64 kb + 32) (hi) and [64 kb + 32, 64 kb + 64) (lo)."""
    L = tc_layout(w.split, w.ksize, w.Cout, w.shuffle)
    full = _tc_matrix(w, L)
    hi = round_tf32(full)
    lo = round_tf32(full - hi)
    nk = L["ktc"] // TC_CHUNK
    out = torch.stack([hi.reshape(L["n_pad"], nk, TC_CHUNK), lo.reshape(L["n_pad"], nk, TC_CHUNK)], 2)
    return out.reshape(L["n_pad"], 2 * L["ktc"]).contiguous()


TC_CHUNK_F16 = 64      # fp16 elements per K step (one 128-byte swizzle row)


def pack_tc_f16(w: PackedGemm) -> torch.Tensor:
    """fp16 operand of the kind::f16 path: the pack_tc layout with 64-channel chunks, [N_pad][K_tc16] K-major, rounded to fp16
    (round-to-nearest-even).  Raises if a weight is outside the fp16 range."""
    L = tc_layout(w.split, w.ksize, w.Cout, w.shuffle)
    chunks = [-(-c // TC_CHUNK_F16) for c in w.split]
    ktc = w.ksize * w.ksize * sum(chunks) * TC_CHUNK_F16
    taps, ctot = w.ksize * w.ksize, sum(w.split)
    n_true = 4 * w.Cout if w.shuffle else w.Cout
    src = w.w32[:, :n_true].reshape(taps, ctot, n_true)
    parts, c0 = [], 0
    for c, ch in zip(w.split, chunks):
        blk = src[:, c0 : c0 + c]
        if ch * TC_CHUNK_F16 != c:
            blk = torch.nn.functional.pad(blk, (0, 0, 0, ch * TC_CHUNK_F16 - c))
        parts.append(blk)
        c0 += c
    kmat = torch.cat(parts, 1).reshape(ktc, n_true)
    out = torch.zeros(L["n_pad"], ktc, dtype=torch.float32, device=w.w32.device)
    if w.shuffle:
        for q in range(4):
            out[q * L["cq_pad"] : q * L["cq_pad"] + w.Cout] = kmat[:, q * w.Cout : (q + 1) * w.Cout].t()
    else:
        out[: w.Cout] = kmat.t()
    if float(out.abs().max()) > 6.0e4:
        raise ValueError(f"{w.name}: weight magnitude {float(out.abs().max()):.3g} does not fit fp16; use precision 'tf32'")
    return out.to(torch.float16).contiguous()
