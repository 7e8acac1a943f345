"""Drop-in for the reference's network/attention.py: ``ATMFormer`` and ``RefineBottleneck``.

Both are parameter-owning shells with the reference's constructor signature and state-dict names
(attention.py:216-263, 393-431); ``forward`` runs the sm_100a block pipeline of atmvfi/engine.py
(window gather + LayerNorm, qkv GEMM, fused window attention / attention-to-motion, projection with
window reverse, Mlp with depth-wise conv).  Inside ``Network.forward`` the same kernels are driven from
the network-level plan; the standalone ``forward`` here exists for callers and tests that use a block
on its own, like the reference's ``__main__`` smoke block (attention.py:501-534).
"""
from __future__ import annotations

import os
import sys

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
if os.path.dirname(_HERE) not in sys.path:
    sys.path.insert(0, os.path.dirname(_HERE))

from atmvfi import _lib                                   # noqa: E402
from atmvfi.arch import NUM_HEADS                         # noqa: E402
from atmvfi.engine import PackedModel, _Block, transformer_block   # noqa: E402
from atmvfi import pack                                   # noqa: E402
from atmvfi.modules import ParamTree, relative_coord_buffer   # noqa: E402
from atmvfi.ops import CudaOps, Map, WinGeom              # noqa: E402


def _to_2tuple(x):
    return tuple(x) if isinstance(x, (tuple, list)) else (x, x)


def _block_schema(dim: int, hidden: int, heads: int, ws):
    S = {}
    ln = lambda n: S.update({n + ".weight": ((dim,), "ln_w"), n + ".bias": ((dim,), "ln_b")})
    lin = lambda n, ci, co, b=True: S.update({n + ".weight": ((co, ci), "linear"), **({n + ".bias": ((co,), "bias_tf")} if b else {})})
    ln("norm1")
    if ws is None:
        lin("attn.qkv", dim, 3 * dim, False)
    else:
        S["attn.relative_coord"] = ((1, 1, 2, ws * ws, ws * ws), "buffer")
        lin("attn.q", dim, dim, False)
        lin("attn.kv", dim, 2 * dim, False)
    lin("attn.proj", dim, dim)
    if ws is not None:
        lin("attn.mlp.0", heads, heads // 2)
        lin("attn.mlp.2", heads // 2, 1)
    ln("norm2")
    lin("mlp.fc1", dim, hidden)
    S["mlp.dwconv.dwconv.weight"] = ((hidden, 1, 3, 3), "dw")
    S["mlp.dwconv.dwconv.bias"] = ((hidden,), "bias_tf")
    lin("mlp.fc2", hidden, dim)
    return S


class _WindowBlock(ParamTree):
    _ATM = False

    def __init__(self, dim, window_size=8, shift_size=0, patch_size=1, num_heads=8, mlp_ratio=4., bidirectional=True,
                 qkv_bias=False, qk_scale=None, drop=0., attn_drop=0., drop_path=0., act_layer=None, norm_layer=None):
        super().__init__()
        assert dim % num_heads == 0, f"dim {dim} should be divided by num_heads {num_heads}."
        if qkv_bias or qk_scale is not None or drop or attn_drop or drop_path:
            raise NotImplementedError("the B200 block implements the configuration the reference networks use: "
                                      "no qkv bias, default scale, no dropout / drop-path")
        if num_heads != NUM_HEADS:
            raise NotImplementedError(f"num_heads must be {NUM_HEADS} (head-mix MLP and kernels are built for it)")
        self.dim, self.num_heads, self.patch_size = dim, num_heads, patch_size
        self.window_size, self.shift_size = _to_2tuple(window_size), _to_2tuple(shift_size)
        self.bidirectional = bidirectional
        ws = self.window_size[0] * patch_size
        self.populate(_block_schema(dim, int(dim * mlp_ratio), num_heads, ws if self._ATM else None))
        self._packed = None
        self._sig = None

    def _set_window_size_(self, window_size, shift_size=0):
        self.window_size, self.shift_size = _to_2tuple(window_size), _to_2tuple(shift_size)
        if self._ATM:
            attn = self._modules["attn"]
            attn.relative_coord = relative_coord_buffer(self.window_size[0] * self.patch_size).to(attn.relative_coord.device)

    def _engine(self, device):
        sig = tuple((n, t.data_ptr(), t._version) for n, t in list(self.named_parameters()) + list(self.named_buffers()))
        if sig != self._sig:
            P = {"blk." + k: v.detach() for k, v in self.state_dict().items()}
            b = _Block()
            f32 = lambda n: P["blk." + n].float().contiguous()
            b.g1, b.b1, b.g2, b.b2 = f32("norm1.weight"), f32("norm1.bias"), f32("norm2.weight"), f32("norm2.bias")
            b.qkv = pack.pack_linear(P, ["blk.attn.q", "blk.attn.kv"] if self._ATM else ["blk.attn.qkv"])
            b.proj = pack.pack_linear(P, ["blk.attn.proj"])
            b.fc1, b.fc2 = pack.pack_linear(P, ["blk.mlp.fc1"]), pack.pack_linear(P, ["blk.mlp.fc2"])
            b.dw_w, b.dw_b = pack.pack_dw(P, "blk.mlp.dwconv.dwconv")
            b.atm = self._ATM
            if self._ATM:
                rc = f32("attn.relative_coord")
                b.rc = rc.reshape(2, rc.shape[-2], rc.shape[-1]).contiguous()
                b.rc_closed = False
                b.mix = (f32("attn.mlp.0.weight"), f32("attn.mlp.0.bias"), f32("attn.mlp.2.weight").reshape(-1).contiguous(), f32("attn.mlp.2.bias"))
            self._packed, self._sig = b, sig
        return CudaOps(device, _lib.FP32), self._packed

    def _run(self, x: torch.Tensor, want_motion: bool):
        if x.dim() != 4:
            raise RuntimeError(f"expected [B,H,W,C] tokens (the layout the reference networks pass), got {tuple(x.shape)}")
        B2, H, W, C = x.shape
        assert C == self.dim
        ops, blk = self._engine(x.device)
        with torch.cuda.device(x.device):
            tok = Map(x.detach().float().contiguous())
            g = WinGeom(B2, H, W, self.window_size[0], self.shift_size[0])
            motion = ops.new_map(B2 // 2, H, W, 4) if want_motion else None
            out = transformer_block(ops, blk, tok, g, motion)
            feat = out.view().reshape(B2, H * W, C)
            if not want_motion:
                return feat
            mv = motion.view()
            return feat, torch.cat([mv[..., 0:2], mv[..., 2:4]], 0).reshape(B2, H * W, 2)


class ATMFormer(_WindowBlock):
    """Cross-frame window attention block that also emits a motion vector per token (attention.py:216-334)."""
    _ATM = True

    def __init__(self, dim, window_size=7, shift_size=0, **kw):
        super().__init__(dim, window_size, shift_size, **kw)

    def forward(self, x, H, W, B):
        """x: [2B, H, W, C] (frame-0 images first) -> (tokens [2B, HW, C], motion [2B, HW, 2])."""
        assert x.shape[0] == 2 * B and x.shape[1] == H and x.shape[2] == W
        return self._run(x, True)


class RefineBottleneck(_WindowBlock):
    """Plain (self-attention) Swin block (attention.py:393-495).  x: [B, H, W, C] -> [B, HW, C]."""
    _ATM = False

    def forward(self, x):
        return self._run(x, False)
