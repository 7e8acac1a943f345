"""Width tables and the parameter schema of the two ATM-VFI networks.

Derived from the reference constructors (network/network_base.py:88-260, network/network_lite.py:88-250):
the two models share one topology and differ only in the numbers below.  ``param_schema`` lists the
reference's 236 state-dict entries so that released checkpoints load with strict=True.
"""
from __future__ import annotations

from collections import OrderedDict
from dataclasses import dataclass
from typing import Tuple

NUM_HEADS = 8
ENHANCE_WINDOW = 8          # feat_enhance_transformer window, fixed (network_base.py:122)
MOTION_OUT = 5              # flow0.xy, flow1.xy, occlusion logit


@dataclass(frozen=True)
class Arch:
    name: str
    enc: Tuple[int, int, int, int]     # encoder widths at 1, 1/2, 1/4, 1/8
    mlp_ratio: int                     # transformer Mlp hidden / dim
    motion_ratio: float                # local motion head hidden / (2 * fused dim)
    last_extra: int                    # extra width of the 1/16 encoder stage
    gmlp_hidden: int                   # global motion head hidden
    refine: int                        # refinement U-Net base width

    @property
    def C(self) -> int:                # fused token width at 1/8
        return self.enc[3] + self.enc[2] + 2 * self.enc[1]

    @property
    def hidden(self) -> int:
        return self.C * self.mlp_ratio

    @property
    def last(self) -> int:
        return self.enc[3] + self.last_extra

    @property
    def GC(self) -> int:               # fused token width at 1/16
        return self.last + self.enc[3] + 2 * self.enc[2]

    @property
    def ghidden(self) -> int:
        return self.GC * self.mlp_ratio

    @property
    def motion_hidden(self) -> int:
        return int(2 * self.C * self.motion_ratio)

    @property
    def dec(self) -> Tuple[int, int, int, int]:   # decoder channel counts incl. the 5 motion channels
        c = self.C
        return (2 * c + MOTION_OUT, c + MOTION_OUT, c // 2 + MOTION_OUT, c // 4 + MOTION_OUT)


BASE = Arch("base", (24, 48, 96, 192), 4, 0.75, 96, 768, 64)
_LITE_GC = (96 + 32) + 96 + 2 * 64
LITE = Arch("lite", (16, 32, 64, 96), 2, 0.5, 32, int(_LITE_GC * 2 * 0.5), 32)
ARCHS = {"base": BASE, "lite": LITE}


def param_schema(a: Arch, local_ws: int = 8, global_ws: int = 12) -> "OrderedDict[str, tuple]":
    """name -> (shape, kind); kind in conv/deconv/dw/linear/bias/prelu/ln_w/ln_b/buffer."""
    S: "OrderedDict[str, tuple]" = OrderedDict()

    def conv(n, ci, co, k=3, tf=False):
        S[n + ".weight"] = ((co, ci, k, k), "conv_tf" if tf else "conv")
        S[n + ".bias"] = ((co,), "bias_tf" if tf else "bias")

    def convp(n, ci, co):
        conv(n + ".0", ci, co)
        S[n + ".1.weight"] = ((co,), "prelu")

    def deconvp(n, ci, co):
        S[n + ".0.weight"] = ((ci, co, 2, 2), "deconv")
        S[n + ".0.bias"] = ((co,), "bias")
        S[n + ".1.weight"] = ((co,), "prelu")

    def norm(n, c):
        S[n + ".weight"] = ((c,), "ln_w")
        S[n + ".bias"] = ((c,), "ln_b")

    def lin(n, ci, co, bias=True):
        S[n + ".weight"] = ((co, ci), "linear")
        if bias:
            S[n + ".bias"] = ((co,), "bias_tf")

    def mlp(n, c, hid):
        lin(n + ".fc1", c, hid)
        S[n + ".dwconv.dwconv.weight"] = ((hid, 1, 3, 3), "dw")
        S[n + ".dwconv.dwconv.bias"] = ((hid,), "bias_tf")
        lin(n + ".fc2", hid, c)

    def fusion(n, fine, mid, coarse):
        conv(n + ".layers.0", mid, mid, tf=True)
        conv(n + ".layers.1", fine, fine, tf=True)
        conv(n + ".layers.2", fine, fine, tf=True)
        cat = mid + 2 * fine + coarse
        conv(n + ".proj", cat, cat, k=1, tf=True)
        norm(n + ".norm", cat)

    def block(n, c, hid, ws):            # ws=None: plain Swin block, else ATMFormer
        norm(n + ".norm1", c)
        if ws is None:
            lin(n + ".attn.qkv", c, 3 * c, bias=False)
            lin(n + ".attn.proj", c, c)
        else:
            S[n + ".attn.relative_coord"] = ((1, 1, 2, ws * ws, ws * ws), "buffer")
            lin(n + ".attn.q", c, c, bias=False)
            lin(n + ".attn.kv", c, 2 * c, bias=False)
            lin(n + ".attn.proj", c, c)
            lin(n + ".attn.mlp.0", NUM_HEADS, NUM_HEADS // 2)
            lin(n + ".attn.mlp.2", NUM_HEADS // 2, 1)
        norm(n + ".norm2", c)
        mlp(n + ".mlp", c, hid)

    widths = (3,) + a.enc
    for i in range(4):
        convp(f"feat_extracts.{i}.0", widths[i], widths[i + 1])
        convp(f"feat_extracts.{i}.1", widths[i + 1], widths[i + 1])
    fusion("cross_scale_feature_fusion", a.enc[1], a.enc[2], a.enc[3])
    for k in range(2):
        block(f"feat_enhance_transformer.{k}", a.C, a.hidden, None)
    for k in range(2):
        block(f"local_motion_atmformer.{k}", a.C, a.hidden, local_ws)
    convp("local_motion_mlp.0", 2 * a.C + NUM_HEADS, a.motion_hidden)
    convp("local_motion_mlp.1", a.motion_hidden, a.motion_hidden)
    conv("local_motion_mlp.2", a.motion_hidden, MOTION_OUT, k=1)
    convp("last_feat_extract.0", a.enc[3], a.last)
    convp("last_feat_extract.1", a.last, a.last)
    fusion("global_feature_fusion", a.enc[2], a.enc[3], a.last)
    for k in range(2):
        block(f"global_motion_atmformer.{k}", a.GC, a.ghidden, global_ws)
    convp("global_motion_mlp.0", 2 * a.GC + NUM_HEADS, a.gmlp_hidden)
    convp("global_motion_mlp.1", a.gmlp_hidden, a.gmlp_hidden)
    conv("global_motion_mlp.2", a.gmlp_hidden, MOTION_OUT, k=1)
    d = a.dec
    for i in range(3):
        n = f"upsample_pyramid.{i}"
        j = 0
        if i > 0:
            S[n + ".0.weight"] = ((d[i],), "prelu")
            j = 1
        deconvp(f"{n}.{j}", d[i], d[i + 1])
        convp(f"{n}.{j + 1}", d[i + 1], d[i + 1])
        conv(f"{n}.{j + 2}", d[i + 1], d[i + 1])
    r = a.refine
    convp("proj", d[3] + 15, r)
    convp("down1.0", r, r)
    convp("down2.0", (d[2] - MOTION_OUT) + r, 2 * r)
    convp("down2.1", 2 * r, 2 * r)
    convp("down3.0", (d[1] - MOTION_OUT) + 2 * r, 4 * r)
    convp("down3.1", 4 * r, 4 * r)
    convp("down3.2", 4 * r, 4 * r)
    deconvp("up1.0", 4 * r, 2 * r)
    convp("up1.1", 2 * r, 2 * r)
    deconvp("up2.0", 4 * r, 2 * r)
    convp("up2.1", 2 * r, r)
    deconvp("up3.0", 2 * r, r)
    convp("refine_head.0", 2 * r, r)
    convp("refine_head.1", r, 3)
    return S
