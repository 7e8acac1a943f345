"""Host orchestration of the ATM-VFI forward on the sm_100a kernels.

``PackedModel`` holds the GEMM operands packed from a reference-schema state-dict.  ``Plan`` is the
launch list for one (batch, height, width, global_motion) shape: all buffers are allocated and all
descriptors are built once; running it is a sequence of C-ABI calls on one stream (optionally replayed
as a CUDA graph).  The dataflow follows network_base.py:433-546 (forward_normal); the layout is
channels-last so tokens [2B, HW, C] and feature maps [2B, H, W, C] are the same bytes and every
cat / slice / einops.rearrange of the reference is a pointer offset.
"""
from __future__ import annotations

import contextlib
from typing import Dict, List, Optional, Tuple

import torch

from . import pack
from .arch import Arch, ENHANCE_WINDOW, MOTION_OUT, NUM_HEADS
from .ops import Map, WinGeom

Params = Dict[str, torch.Tensor]


class _Block:
    pass


def relative_coord_closed_form(ws: int) -> torch.Tensor:
    """[2,N,N] key-minus-query offsets: what attention.py:150-165 stores in the ``relative_coord`` buffer."""
    idx = torch.arange(ws * ws)
    px, py = (idx % ws).float(), (idx // ws).float()
    return torch.stack([px[None, :] - px[:, None], py[None, :] - py[:, None]], 0)


class PackedModel:
    def __init__(self, arch: Arch, P: Params, local_ws: int, global_ws: int, with_global: bool = True):
        self.arch, self.local_ws, self.global_ws = arch, local_ws, global_ws
        a = arch
        f32 = lambda n: P[n].detach().float().contiguous()
        self.enc = [(pack.pack_convp(P, f"feat_extracts.{i}.0"), pack.pack_convp(P, f"feat_extracts.{i}.1")) for i in range(4)]

        def fusion(n, fine, mid, coarse):
            return dict(l0=pack.pack_conv(P, n + ".layers.0"), l1=pack.pack_conv(P, n + ".layers.1"), l2=pack.pack_conv(P, n + ".layers.2"),
                        proj=pack.pack_conv(P, n + ".proj", split=[mid, fine, fine, coarse]),
                        gamma=f32(n + ".norm.weight"), beta=f32(n + ".norm.bias"))

        def block(n, atm):
            b = _Block()
            b.g1, b.b1, b.g2, b.b2 = f32(n + ".norm1.weight"), f32(n + ".norm1.bias"), f32(n + ".norm2.weight"), f32(n + ".norm2.bias")
            b.qkv = pack.pack_linear(P, [n + ".attn.q", n + ".attn.kv"] if atm else [n + ".attn.qkv"])
            b.proj = pack.pack_linear(P, [n + ".attn.proj"])
            b.fc1, b.fc2 = pack.pack_linear(P, [n + ".mlp.fc1"]), pack.pack_linear(P, [n + ".mlp.fc2"])
            b.dw_w, b.dw_b = pack.pack_dw(P, n + ".mlp.dwconv.dwconv")
            b.atm = atm
            if atm:
                rc = f32(n + ".attn.relative_coord")
                b.rc = rc.reshape(2, rc.shape[-2], rc.shape[-1]).contiguous()
                b.rc_closed = bool(torch.equal(b.rc.cpu(), relative_coord_closed_form(int(round(rc.shape[-1] ** 0.5)))))
                b.mix = (f32(n + ".attn.mlp.0.weight"), f32(n + ".attn.mlp.0.bias"),
                         f32(n + ".attn.mlp.2.weight").reshape(-1).contiguous(), f32(n + ".attn.mlp.2.bias"))
            return b

        def head(n, c):
            # inputs: motion of ATMFormer block 0 and block 1 (4 channels each: frame-0 xy, frame-1 xy), tokens of frame 0, frame 1
            return (pack.pack_convp(P, n + ".0", split=[4, 4, c, c]), pack.pack_convp(P, n + ".1"), pack.pack_conv(P, n + ".2"))

        self.fuse_local = fusion("cross_scale_feature_fusion", a.enc[1], a.enc[2], a.enc[3])
        self.enhance = [block(f"feat_enhance_transformer.{k}", False) for k in range(2)]
        self.local_blocks = [block(f"local_motion_atmformer.{k}", True) for k in range(2)]
        self.local_head = head("local_motion_mlp", a.C)
        self.with_global = with_global
        if with_global:
            self.last = (pack.pack_convp(P, "last_feat_extract.0"), pack.pack_convp(P, "last_feat_extract.1"))
            self.fuse_global = fusion("global_feature_fusion", a.enc[2], a.enc[3], a.last)
            self.global_blocks = [block(f"global_motion_atmformer.{k}", True) for k in range(2)]
            self.global_head = head("global_motion_mlp", a.GC)
        d = a.dec
        self.pyramid = []
        for i in range(3):
            n, j = f"upsample_pyramid.{i}", (0 if i == 0 else 1)
            up = pack.pack_deconvp(P, f"{n}.{j}", split=[a.C, a.C, MOTION_OUT] if i == 0 else None)
            c1 = pack.pack_convp(P, f"{n}.{j + 1}")
            c2 = pack.pack_conv(P, f"{n}.{j + 2}")
            nxt = f32(f"upsample_pyramid.{i + 1}.0.weight") if i < 2 else None
            self.pyramid.append((up, c1, c2, nxt))
        r = a.refine
        self.proj = pack.pack_convp(P, "proj", split=[d[3], 15])
        self.down1 = pack.pack_convp(P, "down1.0")
        self.down2 = (pack.pack_convp(P, "down2.0", split=[r, d[2] - MOTION_OUT]), pack.pack_convp(P, "down2.1"))
        self.down3 = (pack.pack_convp(P, "down3.0", split=[2 * r, d[1] - MOTION_OUT]), pack.pack_convp(P, "down3.1"), pack.pack_convp(P, "down3.2"))
        self.up1 = (pack.pack_deconvp(P, "up1.0"), pack.pack_convp(P, "up1.1"))
        self.up2 = (pack.pack_deconvp(P, "up2.0", split=[2 * r, 2 * r]), pack.pack_convp(P, "up2.1"))
        self.up3 = pack.pack_deconvp(P, "up3.0", split=[r, r])
        self.head = (pack.pack_convp(P, "refine_head.0", split=[r, r]), pack.pack_convp(P, "refine_head.1"))


# ------------------------------------------------------------------------------------------------
# building blocks (recorded through ``ops``)
# ------------------------------------------------------------------------------------------------
def transformer_block(ops, blk, tok: Map, g: WinGeom, motion: Optional[Map] = None) -> Map:
    """ATMFormer / RefineBottleneck forward (attention.py:265-334, 433-495) on tokens [B2,H,W,C]."""
    C = tok.C
    hidden = blk.fc1.Cout
    xw = ops.new_win_map(g, C)
    ops.window_gather_ln(tok, xw, g, blk.g1, blk.b1)                       # pad + roll + partition + norm1
    qkv = ops.new_win_map(g, 3 * C, f32=True)       # q | k | v stay fp32 in every mode (the attention MMAs are kind::tf32)
    # head-major q | k | v^T (TMA-fed attention).  Measured on B200 at 1080p: Base (hd 48 / 84) +1.3 % end to end; Lite (hd 28 / 44) was
    # -5 % with the first TMA-fed kernel (its whole third N tile of the qkv linear is the transposed-V store path) and is +1.3 % since the
    # attention kernel requests its operands at the top of the CTA - so every head dim takes it (ATMVFI_QKV_HEADS_MIN_HD restores a floor).
    hm = bool(getattr(ops, "qkv_head_major", False)) and C // NUM_HEADS >= getattr(ops, "qkv_head_major_min_hd", 0)
    ops.gemm_conv([xw], blk.qkv, qkv, act=False, qkv_heads=NUM_HEADS if hm else 0, out_f32=True)
    ao = ops.new_win_map(g, C)
    if blk.atm and motion is not None:
        scratch = torch.empty(g.rows * NUM_HEADS * 2, device=xw.t.device, dtype=torch.float32)
        ops.window_attention(qkv, ao, g, NUM_HEADS, True, blk.rc, blk.mix, motion, 0, scratch,
                             rc_closed_form=getattr(blk, "rc_closed", False), head_major=hm)
    else:
        ops.window_attention(qkv, ao, g, NUM_HEADS, blk.atm, head_major=hm)
    t2 = ops.new_map(tok.B, tok.H, tok.W, C)
    ops.gemm_conv([ao], blk.proj, t2, act=False, residual=xw, win=g)       # proj + residual on the NORMED tokens, window reverse
    t3 = ops.new_map(tok.B, tok.H, tok.W, C)
    ops.layernorm(t2, t3, blk.g2, blk.b2)
    h1 = ops.new_map(tok.B, tok.H, tok.W, hidden)
    ops.gemm_conv([t3.rows()], blk.fc1, h1.rows(), act=False)
    out = ops.new_map(tok.B, tok.H, tok.W, C)
    ok = getattr(ops, "mlp_tail_ok", None)
    if ok is not None and ok(h1, blk.fc2, t2, out):
        # DWConv + GELU produced in shared memory as the A operand of fc2: the activated hidden map never reaches HBM
        ops.mlp_tail(h1, blk.dw_w, blk.dw_b, blk.fc2, t2, out)
        return out
    h2 = ops.new_map(tok.B, tok.H, tok.W, hidden)
    ops.dwconv_gelu(h1, h2, blk.dw_w, blk.dw_b)
    ops.gemm_conv([h2.rows()], blk.fc2, out.rows(), act=False, residual=t2.rows())
    return out


def fusion(ops, f, fine: Map, mid: Map, coarse: Map) -> Map:
    """CrossScaleFeatureFusion.forward (network_base.py:73-85): tokens [2B, H, W, Ccat] after LayerNorm."""
    B, H, W = coarse.B, coarse.H, coarse.W
    y0 = ops.new_map(B, H, W, mid.C)
    ops.gemm_conv([mid], f["l0"], y0, stride=2, act=False)
    y1 = ops.new_map(B, H, W, fine.C)
    ops.gemm_conv([fine], f["l1"], y1, stride=4, dil=1, act=False)
    y2 = ops.new_map(B, H, W, fine.C)
    ops.gemm_conv([fine], f["l2"], y2, stride=4, dil=2, act=False)
    z = ops.new_map(B, H, W, f["proj"].Cout)
    ops.gemm_conv([y0, y1, y2, coarse], f["proj"], z, act=False)
    tok = ops.new_map(B, H, W, z.C)
    ops.layernorm(z, tok, f["gamma"], f["beta"])
    return tok


def motion_branch(ops, blocks, head, tok: Map, ws: int) -> Tuple[Map, Map]:
    """Two ATMFormer blocks (shift 0, ws//2) and the conv motion head (network_base.py:367-415).
    Returns (tokens after the blocks [2B,H,W,C], head [B,H,W,5])."""
    B2, H, W = tok.B, tok.H, tok.W
    B = B2 // 2
    # one 4-channel motion map per block (the two blocks cover different row sets under row slabs)
    motion = [ops.new_map(B, H, W, 4, f32=True), ops.new_map(B, H, W, 4, f32=True)]
    for k, shift in enumerate((0, ws // 2)):
        tok = transformer_block(ops, blocks[k], tok, WinGeom(B2, H, W, ws, shift), motion[k])
    h0, h1, h2 = head
    a = ops.new_map(B, H, W, h0.Cout)
    ops.gemm_conv([ops.to_act(motion[0]), ops.to_act(motion[1]), tok.batch(0, B), tok.batch(B, B)], h0, a)
    b = ops.new_map(B, H, W, h1.Cout)
    ops.gemm_conv([a], h1, b)
    out = ops.new_map(B, H, W, MOTION_OUT, f32=True)      # flows + occlusion logit: fp32 in every mode
    ops.gemm_conv([b], h2, out, act=False, out_f32=True)
    return tok, out


# ------------------------------------------------------------------------------------------------
# liveness-based buffer arena
# ------------------------------------------------------------------------------------------------
_OP_NAMES = frozenset(("window_gather_ln", "gemm_conv", "window_attention", "layernorm", "dwconv_gelu", "conv3x3_first", "copy_map",
                       "nhwc_to_nchw", "nchw_to_nhwc", "resize", "flow_warp_nchw", "flow_warp_nhwc", "l1_mean", "select3", "warp_blend",
                       "pack5_planar", "residual_finish", "pyramid_warp", "mlp_tail"))
_ARENA_ALIGN = 1024


class _DryOps:
    """Stands in for the operator layer during a dry run of ``Plan._build``: buffers are shape-only ("meta") tensors and every
    operator call just notes which buffers it touches, giving each buffer a lifetime [first touch, last touch] in launch order."""

    def __init__(self, real):
        self.real, self.recording, self.step = real, None, 0
        self.bufs: List[list] = []           # [nbytes, first, last, persistent, shape]
        self._idx: Dict[int, int] = {}
        self._keep: List[torch.Tensor] = []

    qkv_head_major = property(lambda s: bool(getattr(s.real, "qkv_head_major", False)))
    qkv_head_major_min_hd = property(lambda s: getattr(s.real, "qkv_head_major_min_hd", 48))
    device = property(lambda s: torch.device("meta"))

    act_f16 = property(lambda s: bool(getattr(s.real, "act_f16", False)))

    def _new(self, shape, dtype=torch.float32) -> torch.Tensor:
        t = torch.empty(tuple(int(x) for x in shape), device="meta", dtype=dtype)
        self._idx[id(t)] = len(self.bufs)
        self._keep.append(t)
        self.bufs.append([t.numel() * t.element_size(), self.step, self.step, False, tuple(t.shape)])
        return t

    def new_map(self, B, H, W, C, zero=False, f32=False) -> Map:
        half = self.act_f16 and not f32
        a = 8 if half else 4
        return Map(self._new((B, H, W, (C + a - 1) // a * a), torch.float16 if half else torch.float32), 0, C)

    def new_win_map(self, g: WinGeom, C, f32=False) -> Map:
        return self.new_map(1, 1, g.rows, C, f32=f32)

    def to_act(self, m: Map) -> Map:
        if not self.act_f16 or m.half:
            return m
        out = self.new_map(m.B, m.H, m.W, m.C)
        self.step += 1
        self._touch([m, out])
        return out

    def new_planar(self, *shape) -> torch.Tensor:
        return self._new(shape)

    def replicated(self):
        return contextlib.nullcontext()

    def mlp_tail_ok(self, *a) -> bool:
        f = getattr(self.real, "mlp_tail_ok", None)
        return bool(f and f(*a))

    def index_of(self, x) -> Optional[int]:
        t = x.t if isinstance(x, Map) else x
        if not isinstance(t, torch.Tensor):
            return None
        base = t._base if t._base is not None else t
        return self._idx.get(id(base))

    def _touch(self, x, persistent=False) -> None:
        if isinstance(x, (list, tuple)):
            for v in x:
                self._touch(v, persistent)
            return
        i = self.index_of(x)
        if i is not None:
            b = self.bufs[i]
            b[2] = self.step
            b[3] = b[3] or persistent

    def __getattr__(self, name):
        if name not in _OP_NAMES or (name == "pyramid_warp" and not hasattr(self.real, name)):
            raise AttributeError(name)

        def op(*args, **kwargs):
            self.step += 1
            keep = name == "copy_map"          # video-stream plans: encoder features cross from one step to the next
            self._touch(args, keep)
            self._touch(list(kwargs.values()), keep)
        return op


def _colour_intervals(bufs) -> Tuple[List[int], int]:
    """Offsets for buffers with lifetimes [first, last] (inclusive, in launch order) such that buffers alive at the same time never
    overlap: largest first, each at the lowest aligned address free during its lifetime.  Returns (offsets, arena bytes)."""
    order = sorted(range(len(bufs)), key=lambda i: -bufs[i][0])
    placed: List[Tuple[int, int, int, int]] = []         # (offset, end, first, last)
    offs = [0] * len(bufs)
    top = 0
    for i in order:
        n, first, last, persistent, _ = bufs[i]
        size = (n + _ARENA_ALIGN - 1) // _ARENA_ALIGN * _ARENA_ALIGN
        if persistent:
            first, last = -1, 1 << 60
        busy = sorted((o, e) for (o, e, f, l) in placed if not (l < first or f > last))
        off = 0
        for o, e in busy:
            if off + size <= o:
                break
            off = max(off, e)
        offs[i] = off
        placed.append((off, off + size, first, last))
        top = max(top, off + size)
    return offs, top


class Plan:
    """Launch list + buffers for one input shape."""

    def __init__(self, ops, model: PackedModel, B: int, H: int, W: int, global_motion: bool, ensemble: bool = False, stream: bool = False):
        ensemble = bool(ensemble and global_motion)      # forward_global_ensemble without global motion is forward_normal's local path
        div = 64 if ensemble else (16 if global_motion else 8)
        if H % div or W % div:
            raise RuntimeError(f"ATM-VFI forward: input {H}x{W} must be a multiple of {div} (global_motion={global_motion}, "
                               f"ensemble={ensemble}); pad with InputPadder as the reference does")
        if ensemble and hasattr(ops, "begin_plan"):
            raise NotImplementedError("the multi-scale global-motion ensemble reduces over whole frames and is not available in row-slab mode")
        if global_motion and not model.with_global:
            raise RuntimeError("global_motion requested but the global-motion weights were not packed")
        if stream and (ensemble or hasattr(ops, "begin_plan")):
            raise NotImplementedError("the video-stream plan (encoder reuse) is not combined with the ensemble or with row slabs")
        self.ops, self.model, self.key = ops, model, (B, H, W, global_motion, ensemble, stream)
        self.stream, self.encode_records, self.encode_graph = stream, None, None
        self.in_use = False                     # video-stream plans are owned by one interpolate_stream generator at a time (runtime.acquire_stream_plan)
        a = model.arch
        self.arena, self.buffer_bytes, self.unshared_bytes = None, None, None
        import os
        use_arena = (not hasattr(ops, "begin_plan") and getattr(ops, "allocator", True) is None and hasattr(ops, "lib")
                     and os.environ.get("ATMVFI_ARENA", "1") != "0")
        if use_arena:
            # Dry run: lifetimes of all buffers in launch order -> interval colouring -> ONE arena in which buffers whose lifetimes do
            # not overlap share addresses (Base 1080p: 19 GB of distinct buffers -> a few GB).  Inputs, the public outputs and
            # whatever a video-stream plan carries to its next step keep private space.  The arena is never zeroed between steps:
            # no kernel reads a lane that its producer did not write in the same step (tests/test_gpu_forward.py poisons it).
            dry = _DryOps(ops)
            dry.recording = []
            self._build(dry, model, a, B, H, W, global_motion, ensemble, stream)
            dry._touch([self.im0, self.im1, list(self.outputs.values()), getattr(self, "ensemble_losses", [])], persistent=True)
            offs, total = _colour_intervals(dry.bufs)
            self.unshared_bytes = sum(b[0] for b in dry.bufs)
            self.buffer_bytes = total
            self.arena = torch.zeros(max(total, _ARENA_ALIGN), dtype=torch.uint8, device=ops.device)
            shapes, cursor = [b[4] for b in dry.bufs], [0]

            def alloc(shape, zero, dtype=torch.float32):
                i = cursor[0]
                cursor[0] += 1
                assert tuple(shape) == shapes[i], f"plan build is not deterministic: buffer {i} is {tuple(shape)}, dry run saw {shapes[i]}"
                n = 2 if dtype == torch.float16 else 4
                for d in shape:
                    n *= int(d)
                assert n == dry.bufs[i][0]
                return self.arena[offs[i] : offs[i] + n].view(dtype).view(*shape)

            ops.allocator = alloc
        ops.recording = rec = []
        try:
            self._build(ops, model, a, B, H, W, global_motion, ensemble, stream)
        finally:
            ops.recording = None
            if use_arena:
                ops.allocator = None
        self.records = rec
        self.graph = None

    # .............................................................................................
    @staticmethod
    def _encode(ops, m: PackedModel, im0: Optional[torch.Tensor], im1: torch.Tensor, B: int, H: int, W: int) -> List[Map]:
        """shared_feat_extraction on the two frames stacked on the batch axis (network_base.py:342-352, 451): 4 levels.
        ``im0 is None``: only the frame-1 half of every level is computed (video-stream plan; the frame-0 half is filled by
        copying the previous pair's frame-1 features)."""
        levels, x = [], None
        sel = (lambda t: t) if im0 is not None else (lambda t: t.batch(B, B))
        for l in range(4):
            c0, c1 = m.enc[l]
            h, w = H >> l, W >> l
            y = ops.new_map(2 * B, h, w, c0.Cout)
            if l == 0:      # 3 -> C0 straight from the planar frames (no channels-last copy of the inputs)
                if im0 is not None:
                    ops.conv3x3_first(im0, c0, y.batch(0, B))
                ops.conv3x3_first(im1, c0, y.batch(B, B))
            else:
                ops.gemm_conv([sel(x)], c0, sel(y), stride=2)
            x = ops.new_map(2 * B, h, w, c1.Cout)
            ops.gemm_conv([sel(y)], c1, sel(x))
            levels.append(x)
        return levels

    @staticmethod
    def _global_head(ops, m: PackedModel, levels: List[Map], B: int, H: int, W: int) -> Map:
        """estimate_global_motion (network_base.py:391-415): 5-channel head [B, H/16, W/16, 5] of an encoder run on HxW frames."""
        h16, w16 = H >> 4, W >> 4
        l0, l1 = m.last
        y = ops.new_map(2 * B, h16, w16, l0.Cout)
        ops.gemm_conv([levels[3]], l0, y, stride=2)
        z = ops.new_map(2 * B, h16, w16, l1.Cout)
        ops.gemm_conv([y], l1, z)
        gtok = fusion(ops, m.fuse_global, levels[2], levels[3], z)
        return motion_branch(ops, m.global_blocks, m.global_head, gtok, m.global_ws)[1]

    def _ensemble_flows(self, ops, m: PackedModel, levels: List[Map], pyr0, pyr1, B: int, H: int, W: int):
        """multiscale_global_motion_ensemble (network_base.py:564-615): global flows estimated at input scales 1, 1/2, 1/4; per
        sample the scale whose flows align the full-resolution frames best wins (device-side select, no host round trip)."""
        P = ops.new_planar
        h16, w16 = H >> 4, W >> 4
        losses, cand0, cand1 = [], [], []
        scratch = P(B, 1, 1, 2048)
        for s in range(3):
            hs, ws_ = H >> s, W >> s
            lev = levels if s == 0 else self._encode(ops, m, pyr0[s], pyr1[s], B, hs, ws_)
            head = self._global_head(ops, m, lev, B, hs, ws_)
            f0, f1 = P(B, 2, hs >> 4, ws_ >> 4), P(B, 2, hs >> 4, ws_ >> 4)
            ops.nhwc_to_nchw(head.chan(0, 2), f0); ops.nhwc_to_nchw(head.chan(2, 2), f1)
            # global_alignmentness (network_base.py:548-562): warp the FULL-resolution frames with the up-scaled flows
            u0, u1 = P(B, 2, H, W), P(B, 2, H, W)
            ops.resize(f0, u0, float(16 << s)); ops.resize(f1, u1, float(16 << s))
            a0, a1 = P(B, 3, H, W), P(B, 3, H, W)
            ops.flow_warp_nchw(self.im0, u0, a0); ops.flow_warp_nchw(self.im1, u1, a1)
            loss = P(B, 1, 1, 1)
            ops.l1_mean(a0, a1, loss, scratch)
            losses.append(loss)
            if s:           # bring the candidate to the 1/16 grid of the original frames (network_base.py:606-611)
                c0, c1 = P(B, 2, h16, w16), P(B, 2, h16, w16)
                ops.resize(f0, c0, float(1 << s)); ops.resize(f1, c1, float(1 << s))
                f0, f1 = c0, c1
            cand0.append(f0); cand1.append(f1)
        g0, g1 = P(B, 2, h16, w16), P(B, 2, h16, w16)
        ops.select3(losses, cand0, g0); ops.select3(losses, cand1, g1)
        self.ensemble_losses = losses
        return g0, g1

    def _build(self, ops, m: PackedModel, a: Arch, B: int, H: int, W: int, glob: bool, ensemble: bool = False, stream: bool = False):
        P = ops.new_planar
        if hasattr(ops, "begin_plan"):          # row-slab mode (slab.SlabOps): partition the rows, open the step
            ops.begin_plan(B, H, W, glob)
        self.im0, self.im1 = P(B, 3, H, W), P(B, 3, H, W)
        pyr0, pyr1 = [self.im0], [self.im1]
        with ops.replicated():        # warp sources: every rank keeps the whole (3-channel) pyramid under row slabs
            for l in range(1, 4):
                pyr0.append(P(B, 3, H >> l, W >> l)); pyr1.append(P(B, 3, H >> l, W >> l))
                ops.resize(pyr0[l - 1], pyr0[l]); ops.resize(pyr1[l - 1], pyr1[l])

        if stream:
            # Video-stream plan: frame k+1 is frame 1 of pair k and frame 0 of pair k+1, so its encoder features are computed once.
            # A step = copy the frame-1 half of levels 1-3 to the frame-0 half (175 MB at 1080p) + encode the new frame 1;
            # `encode_records` alone primes the features of the very first frame (Plan.encode_only).
            rec = ops.recording
            ops.recording = enc = []
            levels = self._encode(ops, m, None, self.im1, B, H, W)
            ops.recording = rec
            for l in (1, 2, 3):
                ops.copy_map(levels[l].batch(B, B), levels[l].batch(0, B))
            rec.extend(enc)
            self.encode_records = enc
        else:
            levels = self._encode(ops, m, self.im0, self.im1, B, H, W)
        tok = fusion(ops, m.fuse_local, levels[1], levels[2], levels[3])     # [2B, H/8, W/8, C]
        h8, w8 = H >> 3, W >> 3

        it_list, w0_list, w1_list = [], [], []
        if glob:
            h16, w16 = H >> 4, W >> 4
            if ensemble:        # forward_global_ensemble (network_base.py:643-649): no 1/16 blend, the lists hold 4 scales
                f0, f1 = self._ensemble_flows(ops, m, levels, pyr0, pyr1, B, H, W)
            else:
                ghead = self._global_head(ops, m, levels, B, H, W)
                i0, i1 = P(B, 3, h16, w16), P(B, 3, h16, w16)
                with ops.replicated():
                    ops.resize(pyr0[3], i0); ops.resize(pyr1[3], i1)
                a0, a1, it = P(B, 3, h16, w16), P(B, 3, h16, w16), P(B, 3, h16, w16)
                f0, f1 = P(B, 2, h16, w16), P(B, 2, h16, w16)
                ops.warp_blend(i0, i1, ghead, a0, a1, it, f0, f1)
                it_list.insert(0, it); w0_list.insert(0, a0); w1_list.insert(0, a1)
            # flows x2 up to 1/8, warp the fused tokens of each frame (network_base.py:471-478)
            f0u, f1u = P(B, 2, h8, w8), P(B, 2, h8, w8)
            with ops.replicated():    # the global flows drive the warps of the whole image pyramid below
                ops.resize(f0, f0u, 2.0); ops.resize(f1, f1u, 2.0)
            fl = ops.new_map(2 * B, h8, w8, 2, f32=True)
            ops.nchw_to_nhwc(f0u, fl.batch(0, B), zero_fill_to=fl.pitch)
            ops.nchw_to_nhwc(f1u, fl.batch(B, B), zero_fill_to=fl.pitch)
            tokw = ops.new_map(2 * B, h8, w8, tok.C)
            ops.flow_warp_nhwc(tok, fl, 0, tokw)
            tok = tokw
            # warp every pyramid level with the global flow, coarse to fine (network_base.py:480-485)
            f0, f1 = f0u, f1u
            # Row slabs: the un-warped pyramid is whole on every rank, so each rank warps only its own rows; the later
            # warp_blend launches read the warped levels in place from their owners.  The flows stay replicated down to 1/2
            # resolution (cheap) so that no up-sampling step needs halo rows from a neighbour.
            if hasattr(ops, "pyramid_warp"):
                # fused: one launch per level warps both frames and (levels 2..0) up-samples the coarser level's flows on the fly
                for l in (3, 2, 1, 0):
                    n0, n1 = P(*pyr0[l].shape), P(*pyr1[l].shape)
                    if l == 3:
                        ops.pyramid_warp(pyr0[l], pyr1[l], f0, f1, False, n0, n1)
                    else:
                        g0, g1 = (P(B, 2, H >> l, W >> l), P(B, 2, H >> l, W >> l)) if l else (None, None)
                        ops.pyramid_warp(pyr0[l], pyr1[l], f0, f1, True, n0, n1, g0, g1)
                        f0, f1 = g0, g1
                    pyr0[l], pyr1[l] = n0, n1
            else:
                for l in (3, 2, 1, 0):
                    n0, n1 = P(*pyr0[l].shape), P(*pyr1[l].shape)
                    ops.flow_warp_nchw(pyr0[l], f0, n0); ops.flow_warp_nchw(pyr1[l], f1, n1)
                    pyr0[l], pyr1[l] = n0, n1
                    if l:
                        g0, g1 = P(B, 2, H >> (l - 1), W >> (l - 1)), P(B, 2, H >> (l - 1), W >> (l - 1))
                        with (ops.replicated() if l > 1 else contextlib.nullcontext()):
                            ops.resize(f0, g0, 2.0); ops.resize(f1, g1, 2.0)
                        f0, f1 = g0, g1

        tok, lhead = motion_branch(ops, m.local_blocks, m.local_head, tok, m.local_ws)
        for k, shift in enumerate((0, ENHANCE_WINDOW // 2)):
            tok = transformer_block(ops, m.enhance[k], tok, WinGeom(2 * B, h8, w8, ENHANCE_WINDOW, shift))

        a0, a1, it = P(B, 3, h8, w8), P(B, 3, h8, w8), P(B, 3, h8, w8)
        ops.warp_blend(pyr0[3], pyr1[3], lhead, a0, a1, it)
        it_list.insert(0, it); w0_list.insert(0, a0); w1_list.insert(0, a1)
        # warp each frame's enhanced features once with the 1/8 flows (network_base.py:504-506)
        fw = ops.new_map(2 * B, h8, w8, tok.C)
        ops.flow_warp_nhwc(tok.batch(0, B), lhead, 0, fw.batch(0, B))
        ops.flow_warp_nhwc(tok.batch(B, B), lhead, 2, fw.batch(B, B))

        srcs = [fw.batch(0, B), fw.batch(B, B), ops.to_act(lhead)]
        f16 = bool(getattr(ops, "act_f16", False))
        skips = []
        flow0 = flow1 = occ1 = occ2 = None
        raw = None
        for i, l in enumerate((2, 1, 0)):
            up, c1, c2, nxt = m.pyramid[i]
            h, w = H >> l, W >> l
            u = ops.new_map(B, h, w, up.Cout)
            ops.gemm_conv(srcs, up, u)
            v = ops.new_map(B, h, w, c1.Cout)
            ops.gemm_conv([u], c1, v)
            raw = ops.new_map(B, h, w, c2.Cout)
            act = ops.new_map(B, h, w, c2.Cout) if nxt is not None else None
            if f16:     # fp16 maps: the level's flows + occlusion logit are also written as an fp32 copy for the warps
                hd = ops.new_map(B, h, w, MOTION_OUT, f32=True)
                ops.gemm_conv([v], c2, raw, act=False, out2=act, prelu2=nxt, head32=hd, head32_c0=raw.C - MOTION_OUT)
            else:
                ops.gemm_conv([v], c2, raw, act=False, out2=act, prelu2=nxt)     # raw level output (+ PReLU'd copy for the next deconv)
                hd = raw.chan(raw.C - MOTION_OUT, MOTION_OUT)
            if l:
                skips.append(raw.chan(0, raw.C - MOTION_OUT))
            a0, a1, it = P(B, 3, h, w), P(B, 3, h, w), P(B, 3, h, w)
            if l == 0:
                flow0, flow1, occ1, occ2 = P(B, 2, h, w), P(B, 2, h, w), P(B, 1, h, w), P(B, 1, h, w)
            ops.warp_blend(pyr0[l], pyr1[l], hd, a0, a1, it, flow0, flow1, occ1, occ2)
            it_list.insert(0, it); w0_list.insert(0, a0); w1_list.insert(0, a1)
            srcs = [act]

        # residual-refinement U-Net (network_base.py:417-431); the ORIGINAL frames go in
        r = a.refine
        imgs = ops.new_map(B, H, W, 15)
        ops.pack5_planar((self.im0, a0, self.im1, a1, it), imgs)
        r0 = ops.new_map(B, H, W, r)
        ops.gemm_conv([raw, imgs], m.proj, r0)
        r1 = ops.new_map(B, H >> 1, W >> 1, r)
        ops.gemm_conv([r0], m.down1, r1, stride=2)
        t = ops.new_map(B, H >> 2, W >> 2, 2 * r)
        ops.gemm_conv([r1, skips.pop()], m.down2[0], t, stride=2)
        r2 = ops.new_map(B, H >> 2, W >> 2, 2 * r)
        ops.gemm_conv([t], m.down2[1], r2)
        t = ops.new_map(B, h8, w8, 4 * r)
        ops.gemm_conv([r2, skips.pop()], m.down3[0], t, stride=2)
        t2 = ops.new_map(B, h8, w8, 4 * r)
        ops.gemm_conv([t], m.down3[1], t2)
        r3 = ops.new_map(B, h8, w8, 4 * r)
        ops.gemm_conv([t2], m.down3[2], r3)
        t = ops.new_map(B, H >> 2, W >> 2, 2 * r)
        ops.gemm_conv([r3], m.up1[0], t)
        u2 = ops.new_map(B, H >> 2, W >> 2, 2 * r)
        ops.gemm_conv([t], m.up1[1], u2)
        t = ops.new_map(B, H >> 1, W >> 1, 2 * r)
        ops.gemm_conv([u2, r2], m.up2[0], t)
        u1 = ops.new_map(B, H >> 1, W >> 1, r)
        ops.gemm_conv([t], m.up2[1], u1)
        u0 = ops.new_map(B, H, W, r)
        ops.gemm_conv([u1, r1], m.up3, u0)
        t = ops.new_map(B, H, W, r)
        ops.gemm_conv([u0, r0], m.head[0], t)
        res = ops.new_map(B, H, W, 3, f32=True)
        ops.gemm_conv([t], m.head[1], res, out_f32=True)
        it_sum, out = P(B, 3, H, W), P(B, 3, H, W)
        ops.residual_finish(res, it, it_sum, out)
        it_list[0] = it_sum      # the reference's in-place ``I_t += residual`` aliases im_t_list[0] (network_base.py:532)

        self.outputs = {"I_t": out, "im_t_list": it_list, "im0_warped_list": w0_list, "im1_warped_list": w1_list,
                        "opt_flow_0": flow0, "opt_flow_1": flow1, "I_t_0": a0, "I_t_1": a1, "occ_mask1": occ1, "occ_mask2": occ2}
        if hasattr(ops, "gather_outputs"):      # row slabs: bring every rank's rows of the public tensors to rank 0
            ops.gather_outputs(self.outputs)

    # .............................................................................................
    def launch(self, stream: Optional[int] = None) -> None:
        """Enqueue every kernel of the forward on ``stream`` (default: torch's current stream)."""
        self.ops.replay(self.records, stream)

    def num_launches(self) -> int:
        return self.ops.count_launches(self.records)

    def encode_only(self, use_graph: bool = False) -> None:
        """Video-stream plan: compute the encoder features of the frame currently in ``self.im1`` (priming the first frame)."""
        assert self.stream, "encode_only belongs to the video-stream plan"
        if not use_graph:
            self.ops.replay(self.encode_records)
            return
        if self.encode_graph is None:
            self.ops.replay(self.encode_records)               # this call's execution; the graph serves later calls
            torch.cuda.current_stream().synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=self._capture_stream()):
                self.ops.replay(self.encode_records)
            self.encode_graph = g
            return
        self.encode_graph.replay()

    def _capture_stream(self) -> torch.cuda.Stream:
        """A capture stream on THIS plan's device (torch.cuda.graph's default capture stream is created once, on whatever device was
        current first: a plan on another GPU would capture nothing and later replay an empty graph)."""
        return torch.cuda.Stream(device=self.ops.device)

    def run(self, im0: torch.Tensor, im1: torch.Tensor, use_graph: bool = False) -> Dict[str, object]:
        self.im0.copy_(im0); self.im1.copy_(im1)
        return self.run_inplace(use_graph)

    def run_inplace(self, use_graph: bool = False) -> Dict[str, object]:
        """Run on whatever ``self.im0`` / ``self.im1`` currently hold (filled by the caller on this stream)."""
        if use_graph:
            if self.graph is None:
                # The first call runs eagerly (lazy module load, attributes) and IS this call's execution; the graph is only
                # captured for the calls that follow.  (Replaying right after the eager run would execute the step twice,
                # which the video-stream step is not idempotent under: its feature copy reads what its encoder overwrites.)
                self.launch()
                torch.cuda.current_stream().synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=self._capture_stream()):
                    self.launch()
                self.graph = g
                return self.outputs
            self.graph.replay()
        else:
            self.launch()
        return self.outputs


def clone_outputs(out: Dict[str, object]) -> Dict[str, object]:
    return {k: ([t.clone() for t in v] if isinstance(v, list) else v.clone()) for k, v in out.items()}
