"""CPU oracle for the ATM-VFI model forward.  TEST INFRASTRUCTURE ONLY.

This file is a functional restatement, in plain torch-CPU fp32 ops, of the reference forward

    /root/reference/network/network_base.py:336-546   (Network.forward / forward_normal)
    /root/reference/network/network_lite.py:329-538   (same graph, narrower widths)
    /root/reference/network/attention.py:8-334,337-495 (ATMFormer, AttentionToMotion, RefineBottleneck)
    /root/reference/network/flow_warp.py:7-60          (flow_warp)
    /root/reference/benchmark/utils.py:57-80           (InputPadder)
    /root/reference/demo_2x.py:54-87                   (inference_2frame host arithmetic)

It is the checker the CUDA path is compared against.  Only ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the product
package (``atm-vfi_b200/``) never does.

Parity status: PINNED.  ``oracle/gen_golden.py`` imports the unmodified reference in the build
container, loads the same deterministic weight sets (``oracle/weights.py``) into it and stores its
outputs under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks this restatement against
those files (the reference itself ships no golden vectors or tests).

The restatement takes a plain ``state_dict`` (name -> tensor, the reference's 236-key schema) and
infers every width from tensor shapes, so one code path serves Base and Lite.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
Params = Dict[str, Tensor]
ENHANCE_WINDOW = 8      # local_motion_args["enhance_window"], network_base.py:122


# --------------------------------------------------------------------------------------------
# small building blocks
# --------------------------------------------------------------------------------------------
def _conv(P: Params, name: str, x: Tensor, stride: int = 1, pad: int = 1, dil: int = 1) -> Tensor:
    return F.conv2d(x, P[name + ".weight"], P[name + ".bias"], stride=stride, padding=pad, dilation=dil)


def _conv_prelu(P: Params, name: str, x: Tensor, stride: int = 1) -> Tensor:
    """``conv()`` helper of the reference: Conv2d(k3,p1) then per-channel PReLU (network_base.py:20-25)."""
    return F.prelu(_conv(P, name + ".0", x, stride=stride), P[name + ".1.weight"])


def _deconv_prelu(P: Params, name: str, x: Tensor) -> Tensor:
    """``deconv()`` helper with kernel 2 / stride 2 / pad 0 (network_base.py:27-32, 202)."""
    y = F.conv_transpose2d(x, P[name + ".0.weight"], P[name + ".0.bias"], stride=2)
    return F.prelu(y, P[name + ".1.weight"])


def _layer_norm(P: Params, name: str, x: Tensor) -> Tensor:
    w = P[name + ".weight"]
    return F.layer_norm(x, (w.numel(),), w, P[name + ".bias"], 1e-5)


def _linear(P: Params, name: str, x: Tensor) -> Tensor:
    return F.linear(x, P[name + ".weight"], P.get(name + ".bias"))


def resize_half(x: Tensor) -> Tensor:
    return F.interpolate(x, scale_factor=0.5, mode="bilinear", align_corners=True)


def upsample_flow2(flow: Tensor) -> Tensor:
    """network_base.py:11-18 with factor 2: bilinear, align_corners=True, values doubled."""
    return F.interpolate(flow, scale_factor=2, mode="bilinear", align_corners=True) * 2


def flow_warp(img: Tensor, flow: Tensor) -> Tensor:
    """flow_warp.py:50-60 -> bilinear_sample (26-47): pixel grid + flow, the lossy fp32
    normalise step, then grid_sample(bilinear, zeros, align_corners=True)."""
    b, _, h, w = img.shape
    assert flow.shape[1] == 2
    ys, xs = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
    base = torch.stack([xs, ys], 0).float()[None].expand(b, -1, -1, -1)
    pos = base + flow
    gx = 2 * pos[:, 0] / (w - 1) - 1
    gy = 2 * pos[:, 1] / (h - 1) - 1
    return F.grid_sample(img, torch.stack([gx, gy], -1), mode="bilinear", padding_mode="zeros", align_corners=True)


# --------------------------------------------------------------------------------------------
# window bookkeeping (attention.py:8-71, 275-305)
# --------------------------------------------------------------------------------------------
def _partition(x: Tensor, ws: int) -> Tensor:
    b, h, w, c = x.shape
    x = x.reshape(b, h // ws, ws, w // ws, ws, c).permute(0, 1, 3, 2, 4, 5)
    return x.reshape(-1, ws * ws, c)


def _unpartition(win: Tensor, ws: int, h: int, w: int) -> Tensor:
    c = win.shape[-1]
    b = win.shape[0] // ((h // ws) * (w // ws))
    x = win.reshape(b, h // ws, w // ws, ws, ws, c).permute(0, 1, 3, 2, 4, 5)
    return x.reshape(b, h, w, c)


def _region_mask(labels: Tensor, ws: int) -> Tensor:
    """labels [1,Hp,Wp,1] -> additive mask [nW,N,N] with -100 where the two tokens' labels differ."""
    lw = _partition(labels, ws).squeeze(-1)
    diff = lw[:, None, :] - lw[:, :, None]
    return torch.where(diff != 0, torch.full_like(diff, -100.0), torch.zeros_like(diff))


def _pad_geometry(h: int, w: int, ws: int) -> Tuple[int, int]:
    return math.ceil(h / ws) * ws - h, math.ceil(w / ws) * ws - w


def _window_prologue(x: Tensor, ws: int, shift: int) -> Tuple[Tensor, Optional[Tensor], Tuple[int, int, int, int]]:
    """centre zero-pad -> roll -> masks.  Returns padded+rolled map, mask, (Hp, Wp, pad_top, pad_left)."""
    _, h, w, _ = x.shape
    ph, pw = _pad_geometry(h, w, ws)
    mask = None
    if ph > 0 or pw > 0:
        lab = torch.zeros(1, h + ph, w + pw, 1)
        rows = (slice(0, ph // 2), slice(ph // 2, h + ph // 2), slice(h + ph // 2, None))
        cols = (slice(0, pw // 2), slice(pw // 2, w + pw // 2), slice(w + pw // 2, None))
        k = 0
        for r in rows:
            for c in cols:
                lab[:, r, c, :] = k
                k += 1
        mask = _region_mask(lab, ws)
        x = F.pad(x, (0, 0, pw // 2, pw - pw // 2, ph // 2, ph - ph // 2))
    hp, wp = h + ph, w + pw
    if shift:
        x = torch.roll(x, shifts=(-shift, -shift), dims=(1, 2))
        lab = torch.zeros(1, hp, wp, 1)
        segs = (slice(0, -ws), slice(-ws, -shift), slice(-shift, None))
        k = 0
        for r in segs:
            for c in segs:
                lab[:, r, c, :] = k
                k += 1
        smask = _region_mask(lab, ws)
        if mask is not None:
            smask = torch.where(mask != 0, torch.full_like(smask, -100.0), smask)
        mask = smask
    return x, mask, (hp, wp, ph // 2, pw // 2)


def _window_epilogue(win: Tensor, ws: int, shift: int, geom: Tuple[int, int, int, int], h: int, w: int) -> Tensor:
    hp, wp, top, left = geom
    x = _unpartition(win, ws, hp, wp)
    if shift:
        x = torch.roll(x, shifts=(shift, shift), dims=(1, 2))
    return x[:, top : top + h, left : left + w, :]


def relative_coord(ws: int) -> Tensor:
    """attention.py:150-165 in closed form: [2, N, N], [0]=x_key-x_query, [1]=y_key-y_query."""
    idx = torch.arange(ws * ws)
    px, py = (idx % ws).float(), (idx // ws).float()
    return torch.stack([px[None, :] - px[:, None], py[None, :] - py[:, None]], 0)


def _softmax_attn(q: Tensor, k: Tensor, mask: Optional[Tensor], heads: int) -> Tensor:
    """q,k: [Bw,N,C] -> attention probabilities [Bw,heads,N,N] (attention.py:189-200)."""
    bw, n, c = q.shape
    hd = c // heads
    qh = q.reshape(bw, n, heads, hd).permute(0, 2, 1, 3)
    kh = k.reshape(bw, n, heads, hd).permute(0, 2, 1, 3)
    logits = (qh @ kh.transpose(-2, -1)) * (hd ** -0.5)
    if mask is not None:
        nw = mask.shape[0]
        logits = (logits.reshape(bw // nw, nw, heads, n, n) + mask[None, :, None]).reshape(bw, heads, n, n)
    return logits.softmax(-1)


def _mlp_dw(P: Params, name: str, x: Tensor, h: int, w: int) -> Tensor:
    """Mlp (attention.py:116-123): fc1 -> depth-wise 3x3 -> exact GELU -> fc2."""
    b, n, _ = x.shape
    y = _linear(P, name + ".fc1", x)
    c = y.shape[-1]
    y = y.transpose(1, 2).reshape(b, c, h, w)
    y = F.conv2d(y, P[name + ".dwconv.dwconv.weight"], P[name + ".dwconv.dwconv.bias"], padding=1, groups=c)
    y = F.gelu(y.reshape(b, c, n).transpose(1, 2))
    return _linear(P, name + ".fc2", y)


def atmformer(P: Params, name: str, x: Tensor, ws: int, shift: int, heads: int = 8) -> Tuple[Tensor, Tensor]:
    """ATMFormer.forward (attention.py:265-334).  x: [2B,H,W,C] (frame-0 batch first).
    Returns tokens [2B,HW,C] and motion [2B,HW,2]."""
    b2, h, w, c = x.shape
    xp, mask, geom = _window_prologue(x, ws, shift)
    xn = _layer_norm(P, name + ".norm1", _partition(xp, ws))
    half = xn.shape[0] // 2
    other = torch.cat([xn[half:], xn[:half]], 0)
    q = _linear(P, name + ".attn.q", xn)
    kv = _linear(P, name + ".attn.kv", other)
    k, v = kv[..., :c], kv[..., c:]
    attn = _softmax_attn(q, k, mask, heads)
    hd = c // heads
    vh = v.reshape(v.shape[0], -1, heads, hd).permute(0, 2, 1, 3)
    app = (attn @ vh).transpose(1, 2).reshape(xn.shape)
    app = _linear(P, name + ".attn.proj", app)
    # attention -> motion: expectation of the key-query offset per head, then the head-mix MLP
    rc = relative_coord(ws)
    if (name + ".attn.relative_coord") in P:
        rc = P[name + ".attn.relative_coord"].reshape(2, ws * ws, ws * ws)
    mot = (attn[:, :, None] * rc[None, None]).sum(-1)            # [Bw, heads, 2, N]
    mot = mot.permute(0, 2, 3, 1)                                # [Bw, 2, N, heads]
    mot = _linear(P, name + ".attn.mlp.2", F.gelu(_linear(P, name + ".attn.mlp.0", mot)))
    mot = mot.squeeze(-1).permute(0, 2, 1)                       # [Bw, N, 2]
    xn = xn + app
    tok = _window_epilogue(xn, ws, shift, geom, h, w).reshape(b2, h * w, c)
    mot = _window_epilogue(mot, ws, shift, geom, h, w).reshape(b2, h * w, 2)
    tok = tok + _mlp_dw(P, name + ".mlp", _layer_norm(P, name + ".norm2", tok), h, w)
    return tok, mot


def swin_block(P: Params, name: str, x: Tensor, ws: int, shift: int, heads: int = 8) -> Tensor:
    """RefineBottleneck.forward (attention.py:433-495): self-attention variant, no motion."""
    b, h, w, c = x.shape
    xp, mask, geom = _window_prologue(x, ws, shift)
    xn = _layer_norm(P, name + ".norm1", _partition(xp, ws))
    qkv = _linear(P, name + ".attn.qkv", xn)
    q, k, v = qkv[..., :c], qkv[..., c : 2 * c], qkv[..., 2 * c :]
    attn = _softmax_attn(q, k, mask, heads)
    hd = c // heads
    vh = v.reshape(v.shape[0], -1, heads, hd).permute(0, 2, 1, 3)
    app = (attn @ vh).transpose(1, 2).reshape(xn.shape)
    xn = xn + _linear(P, name + ".attn.proj", app)
    tok = _window_epilogue(xn, ws, shift, geom, h, w).reshape(b, h * w, c)
    return tok + _mlp_dw(P, name + ".mlp", _layer_norm(P, name + ".norm2", tok), h, w)


# --------------------------------------------------------------------------------------------
# model-level pieces
# --------------------------------------------------------------------------------------------
def cross_scale_fusion(P: Params, name: str, xs: List[Tensor]) -> Tuple[Tensor, int, int]:
    """CrossScaleFeatureFusion.forward with three inputs, fine->coarse (network_base.py:73-85)."""
    fine, mid, coarse = xs
    parts = [
        _conv(P, name + ".layers.0", mid, stride=2, pad=1, dil=1),
        _conv(P, name + ".layers.1", fine, stride=4, pad=1, dil=1),
        _conv(P, name + ".layers.2", fine, stride=4, pad=2, dil=2),
        coarse,
    ]
    y = F.conv2d(torch.cat(parts, 1), P[name + ".proj.weight"], P[name + ".proj.bias"])
    _, _, h, w = y.shape
    return _layer_norm(P, name + ".norm", y.flatten(2).transpose(1, 2)), h, w


def _motion_branch(P: Params, prefix: str, mlp: str, tok: Tensor, h: int, w: int, ws: int):
    """Two ATMFormer blocks (shift 0, ws//2) + the conv motion head.
    tok: [2B,HW,C] -> (flow0, flow1, occ logit->sigmoid, tokens, raw 5-ch head)."""
    b2, _, c = tok.shape
    b = b2 // 2
    motions = []
    for k, shift in enumerate((0, ws // 2)):
        tok, mot = atmformer(P, f"{prefix}.{k}", tok.reshape(b2, h, w, c), ws, shift)
        motions.append(torch.cat([mot[:b], mot[b:]], -1))        # [B, HW, (frame, xy)]
    feat = torch.cat([tok[:b], tok[b:]], -1)                      # [B, HW, 2C]
    head_in = torch.cat(motions + [feat], -1).transpose(1, 2).reshape(b, -1, h, w)
    y = _conv_prelu(P, mlp + ".0", head_in)
    y = _conv_prelu(P, mlp + ".1", y)
    y = _conv(P, mlp + ".2", y, pad=0)
    return y[:, 0:2], y[:, 2:4], torch.sigmoid(y[:, 4:5]), tok, y


def window_sizes(P: Params) -> Tuple[int, int]:
    lw = int(round(math.sqrt(P["local_motion_atmformer.0.attn.relative_coord"].shape[-1])))
    gw = int(round(math.sqrt(P["global_motion_atmformer.0.attn.relative_coord"].shape[-1])))
    return lw, gw


def forward(P: Params, im0: Tensor, im1: Tensor, global_motion: bool = True, ensemble: bool = False) -> Dict[str, object]:
    """forward_normal (network_base.py:433-546) or, with ``ensemble``, forward_global_ensemble (network_base.py:617-712).
    im0, im1: [B,3,H,W] fp32 in [0,1]."""
    with torch.no_grad():
        return _forward(P, im0, im1, global_motion, ensemble)


def _encode(P: Params, x: Tensor):
    """shared_feat_extraction (network_base.py:342-352): returns (1/8 map, [1/2, 1/4, 1/8 maps])."""
    levels = []
    for lvl in range(4):
        x = _conv_prelu(P, f"feat_extracts.{lvl}.0", x, stride=1 if lvl == 0 else 2)
        x = _conv_prelu(P, f"feat_extracts.{lvl}.1", x)
        if lvl:
            levels.append(x)
    return x, levels


def _global_motion(P: Params, x: Tensor, levels: List[Tensor], gws: int):
    """estimate_global_motion (network_base.py:391-415): flows and occlusion mask at 1/16 of the encoder's input."""
    y = _conv_prelu(P, "last_feat_extract.0", x, stride=2)
    y = _conv_prelu(P, "last_feat_extract.1", y)
    gtok, gh, gw = cross_scale_fusion(P, "global_feature_fusion", [levels[1], levels[2], y])
    f0, f1, occ, _, _ = _motion_branch(P, "global_motion_atmformer", "global_motion_mlp", gtok, gh, gw, gws)
    return f0, f1, occ


def upsample_flow(flow: Tensor, factor: int) -> Tensor:
    """upsample_flow (network_base.py:11-18): one bilinear align_corners resize by ``factor``, values scaled by ``factor``."""
    return F.interpolate(flow, scale_factor=factor, mode="bilinear", align_corners=True) * factor


def multiscale_global_motion_ensemble(P: Params, im0: Tensor, im1: Tensor, gws: int):
    """network_base.py:564-615: global flows estimated at input scales 1, 1/2, 1/4; per sample the scale whose flows align the
    two full-resolution frames best (mean L1 of the warped frames, network_base.py:548-562) wins."""
    B = im0.shape[0]
    im = torch.cat([im0, im1], 0)
    cands, losses = [], []
    for scale in range(3):
        if scale:
            im = F.interpolate(im, scale_factor=0.5, mode="bilinear", align_corners=True)
        x, levels = _encode(P, im)
        f0, f1, _ = _global_motion(P, x, levels, gws)
        factor = im0.shape[2] // f0.shape[2]
        a, b = flow_warp(im0, upsample_flow(f0, factor)), flow_warp(im1, upsample_flow(f1, factor))
        losses.append((a - b).abs().mean(dim=[1, 2, 3]))
        cands.append((f0, f1))
    out0, out1 = torch.zeros_like(cands[0][0]), torch.zeros_like(cands[0][1])
    for i in range(B):
        m = min(losses[0][i], losses[1][i], losses[2][i])
        if losses[0][i] == m:
            out0[i], out1[i] = cands[0][0][i], cands[0][1][i]
        elif losses[1][i] == m:
            out0[i], out1[i] = upsample_flow(cands[1][0][i, None], 2)[0], upsample_flow(cands[1][1][i, None], 2)[0]
        else:
            out0[i], out1[i] = upsample_flow(cands[2][0][i, None], 4)[0], upsample_flow(cands[2][1][i, None], 4)[0]
    return out0, out1, torch.stack(losses, 1)


def _forward(P: Params, im0: Tensor, im1: Tensor, global_motion: bool, ensemble: bool = False) -> Dict[str, object]:
    B = im0.shape[0]
    lws, gws = window_sizes(P)
    pyr0, pyr1 = [im0], [im1]
    for _ in range(3):
        pyr0.append(resize_half(pyr0[-1]))
        pyr1.append(resize_half(pyr1[-1]))

    # shared encoder on the two frames stacked on the batch axis (network_base.py:342-352, 451)
    x, levels = _encode(P, torch.cat([im0, im1], 0))
    tok, h, w = cross_scale_fusion(P, "cross_scale_feature_fusion", levels)     # [2B, hw, C] at 1/8
    C = tok.shape[-1]

    it_list, w0_list, w1_list = [], [], []
    if global_motion:
        if ensemble:      # network_base.py:643-649: the 1/16 blend is not produced, the lists hold 4 scales
            f0, f1, _ = multiscale_global_motion_ensemble(P, im0, im1, gws)
        else:
            f0, f1, occ = _global_motion(P, x, levels, gws)
            a, b = flow_warp(resize_half(pyr0[-1]), f0), flow_warp(resize_half(pyr1[-1]), f1)
            it_list.insert(0, occ * a + (1 - occ) * b)
            w0_list.insert(0, a)
            w1_list.insert(0, b)
        f0, f1 = upsample_flow2(f0), upsample_flow2(f1)
        fmap = tok.transpose(1, 2).reshape(2 * B, C, h, w)
        fmap = torch.cat([flow_warp(fmap[:B], f0), flow_warp(fmap[B:], f1)], 0)
        tok = fmap.flatten(2).transpose(1, 2)
        for lvl in (3, 2, 1, 0):
            pyr0[lvl] = flow_warp(pyr0[lvl], f0)
            pyr1[lvl] = flow_warp(pyr1[lvl], f1)
            if lvl:
                f0, f1 = upsample_flow2(f0), upsample_flow2(f1)

    f0, f1, occ, tok, head = _motion_branch(P, "local_motion_atmformer", "local_motion_mlp", tok, h, w, lws)

    # feature enhancement: two plain Swin blocks (network_base.py:354-365)
    # (window fixed at 8: ``enhance_window`` is not touched by __set_local_window_size__)
    for k, shift in enumerate((0, ENHANCE_WINDOW // 2)):
        tok = swin_block(P, f"feat_enhance_transformer.{k}", tok.reshape(2 * B, h, w, C), ENHANCE_WINDOW, shift)
    fmap = tok.transpose(1, 2).reshape(2 * B, C, h, w)

    a, b = flow_warp(pyr0[3], f0), flow_warp(pyr1[3], f1)
    it = occ * a + (1 - occ) * b
    it_list.insert(0, it); w0_list.insert(0, a); w1_list.insert(0, b)

    feat = torch.cat([flow_warp(fmap[:B], f0), flow_warp(fmap[B:], f1), head], 1)
    skips = []
    for i, lvl in enumerate((2, 1, 0)):
        name = f"upsample_pyramid.{i}"
        if i == 0:
            feat = _deconv_prelu(P, name + ".0", feat)
            feat = _conv_prelu(P, name + ".1", feat)
            feat = _conv(P, name + ".2", feat)
        else:
            feat = F.prelu(feat, P[name + ".0.weight"])
            feat = _deconv_prelu(P, name + ".1", feat)
            feat = _conv_prelu(P, name + ".2", feat)
            feat = _conv(P, name + ".3", feat)
        head = feat[:, -5:]
        f0, f1, occ = head[:, 0:2], head[:, 2:4], torch.sigmoid(head[:, 4:5])
        if lvl:
            skips.append(feat[:, :-5])
        a, b = flow_warp(pyr0[lvl], f0), flow_warp(pyr1[lvl], f1)
        it = occ * a + (1 - occ) * b
        it_list.insert(0, it); w0_list.insert(0, a); w1_list.insert(0, b)

    # residual-refinement U-Net (network_base.py:417-431); note: the ORIGINAL frames go in
    r0 = _conv_prelu(P, "proj", torch.cat([feat, im0, a, im1, b, it], 1))
    r1 = _conv_prelu(P, "down1.0", r0, stride=2)
    r2 = _conv_prelu(P, "down2.1", _conv_prelu(P, "down2.0", torch.cat([r1, skips.pop()], 1), stride=2))
    r3 = _conv_prelu(P, "down3.0", torch.cat([r2, skips.pop()], 1), stride=2)
    r3 = _conv_prelu(P, "down3.2", _conv_prelu(P, "down3.1", r3))
    u2 = _conv_prelu(P, "up1.1", _deconv_prelu(P, "up1.0", r3))
    u1 = _conv_prelu(P, "up2.1", _deconv_prelu(P, "up2.0", torch.cat([u2, r2], 1)))
    u0 = _deconv_prelu(P, "up3.0", torch.cat([u1, r1], 1))
    res = _conv_prelu(P, "refine_head.1", _conv_prelu(P, "refine_head.0", torch.cat([u0, r0], 1)))
    # network_base.py:532-533: ``I_t += residual`` is in place, so im_t_list[0] (the same tensor) holds the
    # un-clamped sum, while the returned 'I_t' is the clamped copy made by torch.clamp.
    it_list[0] = it + (2 * torch.sigmoid(res) - 1)
    out = torch.clamp(it_list[0], 0, 1)
    return {
        "I_t": out, "im_t_list": it_list, "im0_warped_list": w0_list, "im1_warped_list": w1_list,
        "opt_flow_0": f0, "opt_flow_1": f1, "I_t_0": a, "I_t_1": b, "occ_mask1": occ, "occ_mask2": 1 - occ,
    }


# --------------------------------------------------------------------------------------------
# host-side arithmetic of inference_2frame (demo_2x.py:54-87) and InputPadder (utils.py:57-80)
# --------------------------------------------------------------------------------------------
def pad_amounts(h: int, w: int, divisor: int = 64) -> Tuple[int, int, int, int]:
    ph = (((h // divisor) + 1) * divisor - h) % divisor
    pw = (((w // divisor) + 1) * divisor - w) % divisor
    return pw // 2, pw - pw // 2, ph // 2, ph - ph // 2      # left, right, top, bottom


def inference_2frame(P: Params, img0, img1, global_motion: bool = True, is_bgr: bool = True):
    """numpy HxWx3 uint8 in, numpy HxWx3 uint8 out; CPU version of demo_2x.inference_2frame."""
    import numpy as np

    if is_bgr:
        img0, img1 = img0[:, :, ::-1].copy(), img1[:, :, ::-1].copy()
    t0 = (torch.tensor(img0.transpose(2, 0, 1)) / 255.0).unsqueeze(0)
    t1 = (torch.tensor(img1.transpose(2, 0, 1)) / 255.0).unsqueeze(0)
    l, r, t, b = pad_amounts(t0.shape[-2], t0.shape[-1])
    t0 = F.pad(t0, (l, r, t, b), mode="replicate")
    t1 = F.pad(t1, (l, r, t, b), mode="replicate")
    pred = forward(P, t0, t1, global_motion)["I_t"][0]
    pred = pred[:, t : pred.shape[-2] - b, l : pred.shape[-1] - r]
    out = np.round(pred.numpy().transpose(1, 2, 0) * 255).astype(np.uint8)
    return out[:, :, ::-1].copy() if is_bgr else out
