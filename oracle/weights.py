"""Deterministic weight sets for parity testing.  TEST INFRASTRUCTURE ONLY.

Released checkpoints live on Google Drive (reference README.md:65-71) and cannot be fetched, so every
parity case uses synthetic weights.  To make them identical in the build container (where the real
reference is run to produce ``tests/golden``) and on the GPU box (where only this repo exists), each
tensor is drawn from its own ``torch.Generator`` seeded by crc32(name): values depend on (name, shape,
seed) only, never on module construction order or torch's global RNG stream.

Two variants:
  * ``default`` - the reference's init statistics (kaiming-uniform-like convs, N(0,0.02) linears, zero
    biases / unit LayerNorm in the transformer parts).  Flows stay ~0 and softmax ~uniform: use it for
    precision tolerances.
  * ``stress``  - non-zero biases everywhere, sharpened attention logits, motion heads scaled so flows
    reach several pixels and cross borders, consumers of those channels compensated.  Use it with loose
    thresholds to catch structural bugs (indexing, masks, borders).

The schema (names and shapes of the reference's 236-entry state-dict) is rebuilt here
from the width tables of network_base.py:92,113-260 and network_lite.py:92-250 and is itself pinned by
``tests/golden/schema_{base,lite}.json`` dumped from the real reference.
"""
from __future__ import annotations

import math
import zlib
from collections import OrderedDict
from typing import Dict, Tuple

import torch

WIDTHS = {
    "base": dict(enc=[24, 48, 96, 192], mlp_ratio=4, motion_ratio=0.75, last_extra=96, gmlp_hidden=768, refine=64),
    "lite": dict(enc=[16, 32, 64, 96], mlp_ratio=2, motion_ratio=0.5, last_extra=32, gmlp_hidden=None, refine=32),
}


def dims(kind: str) -> Dict[str, int]:
    w = WIDTHS[kind]
    e = w["enc"]
    c = e[3] + e[2] + 2 * e[1]
    last = e[3] + w["last_extra"]
    gc = last + e[3] + 2 * e[2]
    return dict(
        e0=e[0], e1=e[1], e2=e[2], e3=e[3], C=c, hidden=int(c * w["mlp_ratio"]), fused=2 * c,
        motion_hidden=int(2 * c * w["motion_ratio"]), last=last, GC=gc, ghidden=int(gc * w["mlp_ratio"]),
        gmlp_hidden=w["gmlp_hidden"] or int(gc * 2 * 0.5), d1=c, d2=c // 2, d3=c // 4, r=w["refine"],
    )


def schema(kind: str, local_ws: int = 8, global_ws: int = 12) -> "OrderedDict[str, Tuple[int, ...]]":
    d = dims(kind)
    S: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()

    def conv(n, ci, co, k=3):
        S[n + ".weight"] = (co, ci, k, k); S[n + ".bias"] = (co,)

    def convp(n, ci, co):
        conv(n + ".0", ci, co); S[n + ".1.weight"] = (co,)

    def deconvp(n, ci, co):
        S[n + ".0.weight"] = (ci, co, 2, 2); S[n + ".0.bias"] = (co,); S[n + ".1.weight"] = (co,)

    def norm(n, c):
        S[n + ".weight"] = (c,); S[n + ".bias"] = (c,)

    def lin(n, ci, co, bias=True):
        S[n + ".weight"] = (co, ci)
        if bias:
            S[n + ".bias"] = (co,)

    def mlp(n, c, hid):
        lin(n + ".fc1", c, hid)
        S[n + ".dwconv.dwconv.weight"] = (hid, 1, 3, 3); S[n + ".dwconv.dwconv.bias"] = (hid,)
        lin(n + ".fc2", hid, c)

    def fusion(n, fine, mid, coarse):
        conv(n + ".layers.0", mid, mid); conv(n + ".layers.1", fine, fine); conv(n + ".layers.2", fine, fine)
        cat = mid + 2 * fine + coarse
        conv(n + ".proj", cat, cat, k=1); norm(n + ".norm", cat)

    def swin(n, c, hid):
        norm(n + ".norm1", c); lin(n + ".attn.qkv", c, 3 * c, bias=False); lin(n + ".attn.proj", c, c)
        norm(n + ".norm2", c); mlp(n + ".mlp", c, hid)

    def atm(n, c, hid, ws):
        norm(n + ".norm1", c)
        S[n + ".attn.relative_coord"] = (1, 1, 2, ws * ws, ws * ws)
        lin(n + ".attn.q", c, c, bias=False); lin(n + ".attn.kv", c, 2 * c, bias=False); lin(n + ".attn.proj", c, c)
        lin(n + ".attn.mlp.0", 8, 4); lin(n + ".attn.mlp.2", 4, 1)
        norm(n + ".norm2", c); mlp(n + ".mlp", c, hid)

    enc = [3, d["e0"], d["e1"], d["e2"], d["e3"]]
    for i in range(4):
        convp(f"feat_extracts.{i}.0", enc[i], enc[i + 1]); convp(f"feat_extracts.{i}.1", enc[i + 1], enc[i + 1])
    fusion("cross_scale_feature_fusion", d["e1"], d["e2"], d["e3"])
    for k in range(2):
        swin(f"feat_enhance_transformer.{k}", d["C"], d["hidden"])
    for k in range(2):
        atm(f"local_motion_atmformer.{k}", d["C"], d["hidden"], local_ws)
    convp("local_motion_mlp.0", d["fused"] + 8, d["motion_hidden"])
    convp("local_motion_mlp.1", d["motion_hidden"], d["motion_hidden"])
    conv("local_motion_mlp.2", d["motion_hidden"], 5, k=1)
    convp("last_feat_extract.0", d["e3"], d["last"]); convp("last_feat_extract.1", d["last"], d["last"])
    fusion("global_feature_fusion", d["e2"], d["e3"], d["last"])
    for k in range(2):
        atm(f"global_motion_atmformer.{k}", d["GC"], d["ghidden"], global_ws)
    convp("global_motion_mlp.0", 2 * d["GC"] + 8, d["gmlp_hidden"])
    convp("global_motion_mlp.1", d["gmlp_hidden"], d["gmlp_hidden"])
    conv("global_motion_mlp.2", d["gmlp_hidden"], 5, k=1)
    chans = [d["fused"] + 5, d["d1"] + 5, d["d2"] + 5, d["d3"] + 5]
    for i in range(3):
        n, ci, co = f"upsample_pyramid.{i}", chans[i], chans[i + 1]
        if i == 0:
            deconvp(n + ".0", ci, co); convp(n + ".1", co, co); conv(n + ".2", co, co)
        else:
            S[n + ".0.weight"] = (ci,); deconvp(n + ".1", ci, co); convp(n + ".2", co, co); conv(n + ".3", co, co)
    r = d["r"]
    convp("proj", d["d3"] + 5 + 15, r)
    convp("down1.0", r, r)
    convp("down2.0", d["d2"] + r, 2 * r); convp("down2.1", 2 * r, 2 * r)
    convp("down3.0", d["d1"] + 2 * r, 4 * r); convp("down3.1", 4 * r, 4 * r); convp("down3.2", 4 * r, 4 * r)
    deconvp("up1.0", 4 * r, 2 * r); convp("up1.1", 2 * r, 2 * r)
    deconvp("up2.0", 4 * r, 2 * r); convp("up2.1", 2 * r, r)
    deconvp("up3.0", 2 * r, r)
    convp("refine_head.0", 2 * r, r); convp("refine_head.1", r, 3)
    return S


def _randn(name: str, shape, seed: int) -> torch.Tensor:
    g = torch.Generator()
    g.manual_seed((zlib.crc32(name.encode()) + 1000003 * seed) & 0x7FFFFFFF)
    return torch.randn(tuple(shape), generator=g, dtype=torch.float32)


def _relative_coord(ws: int) -> torch.Tensor:
    idx = torch.arange(ws * ws)
    px, py = (idx % ws).float(), (idx // ws).float()
    return torch.stack([px[None, :] - px[:, None], py[None, :] - py[:, None]], 0)[None, None].contiguous()


_TRANSFORMER_PARTS = ("cross_scale_feature_fusion", "global_feature_fusion", "feat_enhance_transformer",
                      "local_motion_atmformer", "global_motion_atmformer")


def make_weights(kind: str, variant: str = "default", seed: int = 0, local_ws: int = 8, global_ws: int = 12) -> Dict[str, torch.Tensor]:
    assert variant in ("default", "stress", "ensemble", "varflow")
    S = schema(kind, local_ws, global_ws)
    P: Dict[str, torch.Tensor] = OrderedDict()
    for name, shape in S.items():
        if name.endswith("relative_coord"):
            P[name] = _relative_coord(int(round(math.sqrt(shape[-1]))))
            continue
        z = _randn(name, shape, seed)
        in_tf = name.startswith(_TRANSFORMER_PARTS)
        if len(shape) == 4:                                   # conv / deconv / depth-wise weights
            if in_tf:                                         # N(0, sqrt(2/fan_out)), attention.py:109-112
                fan_out = shape[2] * shape[3] * shape[0] // (shape[0] if shape[1] == 1 else 1)
                P[name] = z * math.sqrt(2.0 / fan_out)
            else:                                             # ~ kaiming_uniform(a=sqrt(5)): var = 1/(3 fan_in)
                fan_in = shape[1] * shape[2] * shape[3]
                if ".0.weight" in name and shape[2] == 2:     # ConvTranspose [Cin,Cout,2,2]: torch uses size(1)*k*k
                    fan_in = shape[1] * 4
                P[name] = z * math.sqrt(1.0 / (3.0 * fan_in))
        elif len(shape) == 2:                                 # Linear
            P[name] = z.clamp(-2, 2) * 0.02
        elif "norm" in name and in_tf:                        # LayerNorm
            P[name] = torch.ones(shape) if name.endswith("weight") else torch.zeros(shape)
        elif name.endswith(".bias"):
            if in_tf:
                P[name] = torch.zeros(shape)
            else:
                owner = name[: -len(".bias")] + ".weight"
                ws = S[owner]
                fan_in = (ws[1] * 4) if ws[2] == 2 else ws[1] * ws[2] * ws[3]
                P[name] = (torch.rand(shape, generator=torch.Generator().manual_seed(zlib.crc32(name.encode()) & 0x7FFFFFFF)) * 2 - 1) / math.sqrt(fan_in)
        else:                                                 # PReLU slope
            P[name] = torch.full(shape, 0.25)
    if variant in ("stress", "ensemble", "varflow"):
        _apply_stress(P, kind, seed)
    if variant == "varflow":
        _apply_varflow(P, kind)
    if variant == "ensemble":
        # Scale-selection probe for forward_global_ensemble (network_base.py:564-615): the global head emits a constant flow of
        # +-0.125 grid pixels (plus a small data term), i.e. +-2 / +-4 / +-8 full-resolution pixels when estimated at input
        # scale 1, 1/2, 1/4.  With the "shift" frames below (sample i shifted by 4, 8, 16 px) a different scale wins per sample.
        P["global_motion_mlp.2.weight"] = P["global_motion_mlp.2.weight"] * 0.002
        P["global_motion_mlp.2.bias"] = torch.tensor([0.125, 0.0, -0.125, 0.0, 0.0])
    return P


def _apply_stress(P: Dict[str, torch.Tensor], kind: str, seed: int) -> None:
    """SURVEY.md section 4 recipe, re-expressed on the name-seeded generator."""
    d = dims(kind)
    for name in list(P):
        if name.endswith("relative_coord"):
            continue
        t = P[name]
        if name.endswith(".bias"):
            P[name] = t + 0.02 * _randn(name + "#b", t.shape, seed)
        elif t.dim() == 1 and "norm" in name:
            P[name] = 1.0 + 0.1 * _randn(name + "#g", t.shape, seed)
        elif t.dim() == 1:
            P[name] = 0.25 + 0.05 * _randn(name + "#s", t.shape, seed)
    for name in list(P):
        if name.endswith(("attn.q.weight", "attn.kv.weight", "attn.qkv.weight")):
            P[name] = P[name] * 10.0
        if name.endswith(("attn.mlp.0.weight", "attn.mlp.2.weight")):
            P[name] = P[name] * math.sqrt(30.0)
    g, g_glob, g_occ = 100.0, 5.0, 50.0
    producers = {"local_motion_mlp.2": g, "global_motion_mlp.2": g_glob,
                 "upsample_pyramid.0.2": g, "upsample_pyramid.1.3": g, "upsample_pyramid.2.3": g}
    for n, gain in producers.items():
        gains = torch.tensor([gain, gain, gain, gain, g_occ])
        P[n + ".weight"][-5:] *= gains.view(5, 1, 1, 1)
        P[n + ".bias"][-5:] *= gains
    inv = (1.0 / torch.tensor([g, g, g, g, g_occ])).view(5, 1, 1, 1)
    for n in ("upsample_pyramid.0.0.0.weight", "upsample_pyramid.1.1.0.weight", "upsample_pyramid.2.1.0.weight"):
        P[n][-5:] *= inv                                       # ConvTranspose layout [Cin, Cout, 2, 2]
    P["proj.0.weight"][:, d["d3"] : d["d3"] + 5] *= inv.view(1, 5, 1, 1)


def _apply_varflow(P: Dict[str, torch.Tensor], kind: str) -> None:
    """On top of the stress set: the flow channels of the four motion producers get their WEIGHT rows scaled by a further
    x10 / x10 / x15 / x20 (biases untouched, consumers compensated), so the flows are driven by the features instead of by the
    biases.  Measured with the oracle (Base, 256x448 texture frames): final flows with a spatial std of 2-3.5 px and a range of
    45 px per component, means of -49 / +10 px (warps cross the frame border over a 50-pixel margin); the plain stress set gives
    flows that are constant to 0.15 px.  Rounding noise is amplified by the same gains: structural checks only."""
    d = dims(kind)
    pairs = (("local_motion_mlp.2", 10.0, "upsample_pyramid.0.0.0.weight"), ("upsample_pyramid.0.2", 10.0, "upsample_pyramid.1.1.0.weight"),
             ("upsample_pyramid.1.3", 15.0, "upsample_pyramid.2.1.0.weight"), ("upsample_pyramid.2.3", 20.0, "proj.0.weight"))
    for prod, r, cons in pairs:
        P[prod + ".weight"][-5:-1] *= r
        if cons == "proj.0.weight":
            P[cons][:, d["d3"] : d["d3"] + 4] /= r
        else:
            P[cons][-5:-1] /= r


def synthetic_triplet(batch: int, h: int, w: int, seed: int = 1234):
    """(frame 0, ground-truth middle frame, frame 1) of a smooth texture with fine detail that moves by (-2 rows, +3 columns) per
    half step: a clip with a KNOWN middle frame, for PSNR-versus-ground-truth comparisons."""
    g = torch.Generator().manual_seed(seed)
    low = torch.rand(batch, 3, h // 8 + 3, w // 8 + 3, generator=g)
    big = torch.nn.functional.interpolate(low, size=(h + 16, w + 16), mode="bicubic", align_corners=True).clamp(0, 1)
    fine = 0.15 * (torch.rand(batch, 3, h + 16, w + 16, generator=g) - 0.5)
    big = (big + fine).clamp(0, 1)
    crop = lambda dy, dx: big[:, :, dy : dy + h, dx : dx + w].contiguous()
    return crop(8, 4), crop(6, 7), crop(4, 10)


def synthetic_frames(batch: int, h: int, w: int, seed: int = 1234, kind: str = "noise"):
    """Frame pairs in [0,1).  ``noise`` = torch.rand; ``texture`` = a smooth random field and a
    shifted copy, so that flows are meaningful with the stress weights."""
    g = torch.Generator().manual_seed(seed)
    if kind == "shift":       # sample i: frame 1 = frame 0 moved by 4 * 2^(i % 3) pixels along x (smooth texture)
        low = torch.rand(batch, 3, h // 16 + 3, (w + 32) // 16 + 3, generator=g)
        big = torch.nn.functional.interpolate(low, size=(h, w + 32), mode="bicubic", align_corners=True).clamp(0, 1)
        im0 = big[:, :, :, 0:w].contiguous()
        im1 = torch.stack([big[i, :, :, 4 * 2 ** (i % 3) : 4 * 2 ** (i % 3) + w] for i in range(batch)]).contiguous()
        return im0, im1
    if kind == "noise":
        return torch.rand(batch, 3, h, w, generator=g), torch.rand(batch, 3, h, w, generator=g)
    low = torch.rand(batch, 3, h // 8 + 3, w // 8 + 3, generator=g)
    big = torch.nn.functional.interpolate(low, size=(h + 16, w + 16), mode="bicubic", align_corners=True).clamp(0, 1)
    fine = 0.15 * (torch.rand(batch, 3, h + 16, w + 16, generator=g) - 0.5)
    big = (big + fine).clamp(0, 1)
    return big[:, :, 8 : 8 + h, 8 : 8 + w].contiguous(), big[:, :, 5 : 5 + h, 12 : 12 + w].contiguous()
