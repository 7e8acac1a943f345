"""CPU emulation of the C-ABI operator CONTRACTS (include/atmvfi.h) - TEST INFRASTRUCTURE ONLY.

It lets the CPU suite check the host logic of the product (weight packing in pack.py, buffer slicing
and launch order in engine.py) against the oracle without a GPU: each op is re-expressed with stock
torch CPU functions from the header's documentation, NOT from the CUDA sources.  Nothing under
``atm-vfi_b200/`` imports this module; the product path has no CPU fallback.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import torch
import torch.nn.functional as F

from atmvfi.ops import Map, PackedGemm, WinGeom, round_up


def _partition(x, ws):
    b, h, w, c = x.shape
    return x.reshape(b, h // ws, ws, w // ws, ws, c).permute(0, 1, 3, 2, 4, 5).reshape(-1, ws * ws, c)


def _unpartition(win, ws, b, h, w):
    c = win.shape[-1]
    return win.reshape(b, h // ws, w // ws, ws, ws, c).permute(0, 1, 3, 2, 4, 5).reshape(b, h, w, c)


def _win_forward(x, g: WinGeom):
    """[B2,H,W,C] -> window-major rows [rows, C] (centre zero pad, roll, partition)."""
    pt, pl = g.pad_top, g.pad_left
    x = F.pad(x, (0, 0, pl, g.Wp - g.W - pl, pt, g.Hp - g.H - pt))
    if g.shift:
        x = torch.roll(x, (-g.shift, -g.shift), (1, 2))
    return _partition(x, g.ws).reshape(g.rows, -1)


def _win_reverse(rows, g: WinGeom):
    x = _unpartition(rows.reshape(-1, g.ws * g.ws, rows.shape[-1]), g.ws, g.B2, g.Hp, g.Wp)
    if g.shift:
        x = torch.roll(x, (g.shift, g.shift), (1, 2))
    return x[:, g.pad_top : g.pad_top + g.H, g.pad_left : g.pad_left + g.W]


def _labels(g: WinGeom):
    """per-row mask label in the window frame: (pad label on the UN-rolled frame, shift label)."""
    ys = torch.arange(g.Hp)[:, None].expand(g.Hp, g.Wp)
    xs = torch.arange(g.Wp)[None, :].expand(g.Hp, g.Wp)
    lab = torch.zeros(g.Hp, g.Wp, dtype=torch.long)
    if g.Hp != g.H or g.Wp != g.W:
        ly = (ys >= g.pad_top).long() + (ys >= g.pad_top + g.H).long()
        lx = (xs >= g.pad_left).long() + (xs >= g.pad_left + g.W).long()
        lab = ly * 3 + lx
    if g.shift:
        ly = (ys >= g.Hp - g.ws).long() + (ys >= g.Hp - g.shift).long()
        lx = (xs >= g.Wp - g.ws).long() + (xs >= g.Wp - g.shift).long()
        lab = lab * 9 + ly * 3 + lx
    return _partition(lab[None, :, :, None].float(), g.ws).squeeze(-1)      # [nW, N]


def _warp(img, flow):
    b, _, h, w = img.shape
    ys, xs = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
    px, py = xs.float() + flow[:, 0], ys.float() + flow[:, 1]
    grid = torch.stack([2 * px / (w - 1) - 1, 2 * py / (h - 1) - 1], -1)
    return F.grid_sample(img, grid, mode="bilinear", padding_mode="zeros", align_corners=True)


def _put(dst, val, rows, dim=1):
    """dst <- val on the row window [y0, y1) of axis `dim` (all rows when rows is None)."""
    val = val.reshape(dst.shape) if val.shape != dst.shape else val
    if rows is None:
        dst.copy_(val)
    else:
        idx = [slice(None)] * dst.dim()
        idx[dim] = slice(rows[0], rows[1])
        dst[tuple(idx)] = val[tuple(idx)]


def _win_token_mask(g: WinGeom, rows):
    """[B2,H,W] bool: tokens that belong to the window rows [k0, k1) (after pad / roll / partition)."""
    nwy, nwx, N = g.Hp // g.ws, g.Wp // g.ws, g.ws * g.ws
    ind = torch.zeros(g.B2, nwy, nwx, N, 1)
    ind[:, rows[0] : rows[1]] = 1
    return _win_reverse(ind.reshape(g.rows, 1), g)[..., 0] > 0


def _to_heads(y, heads):
    """[R, 3C] q|k|v rows -> flat head-major buffer (include/atmvfi.h ATMVFI_OUT_QKV_HEADS)."""
    R, C3 = y.shape
    C = C3 // 3
    hd = C // heads
    q = y[:, :C].reshape(R, heads, hd).permute(1, 0, 2)
    k = y[:, C : 2 * C].reshape(R, heads, hd).permute(1, 0, 2)
    vt = y[:, 2 * C :].t()
    return torch.cat([q.reshape(-1), k.reshape(-1), vt.reshape(-1)])


def _from_heads(flat, R, C, heads):
    hd = C // heads
    q = flat[: C * R].reshape(heads, R, hd).permute(1, 0, 2).reshape(R, C)
    k = flat[C * R : 2 * C * R].reshape(heads, R, hd).permute(1, 0, 2).reshape(R, C)
    v = flat[2 * C * R :].reshape(C, R).t()
    return torch.cat([q, k, v], 1)


def round_tf32(t):
    """fp32 -> nearest TF32 (ties away from zero): what the producers' cvt.rna.tf32.f32 and pack.pack_tc do."""
    bits = t.contiguous().view(torch.int32)
    return ((bits + 0x1000) & ~0x1FFF).view(torch.float32)


def trunc_tf32(t):
    """fp32 -> TF32 by dropping the low 13 mantissa bits: what tcgen05 kind::tf32 does to its operands."""
    return (t.contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32)


class EmulOps:
    def __init__(self, qkv_head_major: bool = False, tf32: bool = False, round_outputs: bool = False):
        """``tf32``: GEMM-shaped layers with >= 16 input channels see their activations truncated and their weights rounded to
        TF32 (products accumulated exactly), like the tcgen05 path; ``round_outputs``: stored feature maps are rounded to TF32."""
        self.recording: Optional[List] = None
        self.launches = 0
        self.qkv_head_major = qkv_head_major
        self.qkv_head_major_min_hd = 0          # tests exercise the layout on every model size
        self.tf32, self.round_outputs = tf32, round_outputs

    act_f16 = False

    def to_act(self, m):
        return m

    def new_map(self, B, H, W, C, zero=False, f32=False):
        return Map(torch.full((B, H, W, round_up(C, 4)), float("nan") if round_up(C, 4) == C and not zero else 0.0), 0, C)

    def new_win_map(self, g: WinGeom, C, f32=False):
        return self.new_map(1, 1, g.rows, C)

    def new_planar(self, *shape):
        return torch.full(shape, float("nan"))

    def replicated(self):
        import contextlib
        return contextlib.nullcontext()

    def _emit(self, fn):
        if self.recording is not None:
            self.recording.append(fn)
        else:
            fn()

    def emit_host(self, fn):
        self._emit(fn)

    def replay(self, records, stream=None):
        for fn in records:
            fn()

    @staticmethod
    def count_launches(records):
        return len(records)

    # ---------------------------------------------------------------------------------------------
    def gemm_conv(self, srcs: Sequence[Map], w: PackedGemm, out: Map, *, stride=1, dil=1, act=True, residual=None,
                  out2=None, prelu2=None, win: Optional[WinGeom] = None, precision=None, rows=None, qkv_heads=0, out_f32=False):
        assert [s.C for s in srcs] == list(w.split)
        k, ci = w.ksize, sum(w.split)
        n_tot = 4 * w.Cout if w.shuffle else w.Cout
        wk = w.w32[:, :n_tot]
        assert wk.shape[0] == k * k * ci

        tc = self.tf32 and ci >= 16
        rnd = (lambda t: round_tf32(t)) if (self.round_outputs and self.tf32) else (lambda t: t)

        def run():
            x = torch.cat([s.view() for s in srcs], -1)                       # [B,H,W,Ci]
            if tc:          # operands as the tensor core sees them; fp64 accumulation stands in for "exact"
                return run_tc(trunc_tf32(x).double(), round_tf32(wk).double())
            if w.shuffle:
                wt = wk.reshape(ci, 2, 2, w.Cout).permute(0, 3, 1, 2)          # [Ci,Co,2,2]
                y = F.conv_transpose2d(x.permute(0, 3, 1, 2), wt, w.bias, stride=2).permute(0, 2, 3, 1)
            else:
                wt = wk.reshape(k, k, ci, w.Cout).permute(3, 2, 0, 1)          # [Co,Ci,k,k]
                y = F.conv2d(x.permute(0, 3, 1, 2), wt, w.bias, stride=stride, padding=dil * (k - 1) // 2, dilation=dil).permute(0, 2, 3, 1)
            if residual is not None:
                y = y + residual.view().reshape(y.shape)
            if act and w.prelu is not None:
                y = torch.where(y > 0, y, y * w.prelu)
            y2 = torch.where(y > 0, y, y * prelu2) if out2 is not None else None
            if win is not None:
                y = _win_reverse(y.reshape(win.rows, -1), win)
                if rows is None:
                    out.view().copy_(y.reshape(out.view().shape))
                else:      # rows = window rows of the GEMM grid: only their tokens are produced
                    m = _win_token_mask(win, rows)
                    ov = out.view().reshape(y.shape)
                    ov[m] = y[m]
                return
            if qkv_heads:       # head-major q | k | v^T; with a row window only the rows of that window are produced
                R = out.view().numel() // w.Cout
                flat = out.t.reshape(-1)
                new = _to_heads(y.reshape(R, w.Cout), qkv_heads)
                if rows is None:
                    flat.copy_(new)
                else:
                    sel = torch.zeros(y.shape[:3], dtype=torch.bool)
                    sel[:, rows[0] : rows[1]] = True
                    keep = _to_heads(sel.reshape(R, 1).expand(R, w.Cout).float(), qkv_heads) > 0
                    flat[keep] = new[keep]
                return
            orow = rows if (rows is None or not w.shuffle) else (2 * rows[0], 2 * rows[1])
            _put(out.view(), y, orow)
            if out2 is not None:
                _put(out2.view(), y2, orow)

        def run_tc(x, wk):
            """Same contract on TF32 operands: conv in fp64, epilogue (bias, residual, PReLU) in the kernel's order, optional
            TF32 rounding of what is stored.  Plain / transposed / window-reverse / dual-output / head-major layouts."""
            assert rows is None
            bias = None if w.bias is None else w.bias.double()
            if w.shuffle:
                wt = wk.reshape(ci, 2, 2, w.Cout).permute(0, 3, 1, 2)
                y = F.conv_transpose2d(x.permute(0, 3, 1, 2), wt, bias, stride=2).permute(0, 2, 3, 1)
            else:
                wt = wk.reshape(k, k, ci, w.Cout).permute(3, 2, 0, 1)
                y = F.conv2d(x.permute(0, 3, 1, 2), wt, bias, stride=stride, padding=dil * (k - 1) // 2, dilation=dil).permute(0, 2, 3, 1)
            if residual is not None:
                y = y + residual.view().reshape(y.shape).double()
            if act and w.prelu is not None:
                y = torch.where(y > 0, y, y * w.prelu.double())
            y2 = rnd(torch.where(y > 0, y, y * prelu2.double()).float()) if out2 is not None else None
            y = rnd(y.float())
            if win is not None:
                out.view().copy_(_win_reverse(y.reshape(win.rows, -1), win).reshape(out.view().shape))
                return
            if qkv_heads:
                out.t.reshape(-1).copy_(_to_heads(y.reshape(-1, w.Cout), qkv_heads))
                return
            _put(out.view(), y, None)
            if out2 is not None:
                _put(out2.view(), y2, None)

        self._emit(run)

    def conv3x3_first(self, img, w: PackedGemm, out: Map, rows=None):
        def run():
            wt = w.w32[:, : w.Cout].reshape(3, 3, 3, w.Cout).permute(3, 2, 0, 1)
            y = F.conv2d(img, wt, w.bias, padding=1)
            y = torch.where(y > 0, y, y * w.prelu.view(1, -1, 1, 1))
            _put(out.view(), y.permute(0, 2, 3, 1), rows)
        self._emit(run)

    def pack5_planar(self, imgs, out: Map, rows=None):
        def run():
            v = torch.cat([t.permute(0, 2, 3, 1) for t in imgs] + [torch.zeros_like(imgs[0][:, :1]).permute(0, 2, 3, 1)], -1)
            _put(out.t[..., :16], v, rows)
        self._emit(run)

    def layernorm(self, x: Map, out: Map, gamma, beta, rows=None):
        self._emit(lambda: _put(out.view(), F.layer_norm(x.view(), (x.C,), gamma, beta, 1e-5), rows))

    def window_gather_ln(self, tok: Map, win: Map, g: WinGeom, gamma, beta, rows=None):
        def run():
            wr = _win_forward(tok.view().reshape(g.B2, g.H, g.W, tok.C), g)
            y = F.layer_norm(wr, (tok.C,), gamma, beta, 1e-5)
            shp = (g.B2, g.Hp // g.ws, g.ws * g.Wp, tok.C)       # window-major rows seen as [image][window row][tokens]
            _put(win.view().reshape(shp), y.reshape(shp), rows)
        self._emit(run)

    def window_attention(self, qkv: Map, out: Map, g: WinGeom, heads, cross, rc=None, mix=None, motion=None, motion_off=0, scratch=None, rc_closed_form=False, rows=None,
                         head_major=False):
        def run():
            C = out.C
            N = g.ws * g.ws
            hd = C // heads
            x = qkv.view().reshape(-1, N, 3 * C)
            if head_major:
                x = _from_heads(qkv.t.reshape(-1), g.rows, C, heads).reshape(-1, N, 3 * C)
            q, k, v = x[..., :C], x[..., C : 2 * C], x[..., 2 * C :]
            if cross:
                half = x.shape[0] // 2
                k, v = torch.cat([k[half:], k[:half]]), torch.cat([v[half:], v[:half]])
            sp = lambda t: t.reshape(-1, N, heads, hd).permute(0, 2, 1, 3)
            logits = (sp(q) @ sp(k).transpose(-1, -2)) * (hd ** -0.5)
            if g.shift or g.Hp != g.H or g.Wp != g.W:
                lab = _labels(g)                                               # [nW, N]
                m = (lab[:, :, None] != lab[:, None, :]).float() * -100.0      # [nW, N, N]
                nW = m.shape[0]
                logits = (logits.reshape(-1, nW, heads, N, N) + m[None, :, None]).reshape(-1, heads, N, N)
            p = logits.softmax(-1)
            shp = (g.B2, g.Hp // g.ws, g.ws * g.Wp, C)
            _put(out.view().reshape(shp), (p @ sp(v)).transpose(1, 2).reshape(shp), rows)
            if motion is not None:
                mo = (p[:, :, None] * rc[None, None]).sum(-1).permute(0, 2, 3, 1)          # [Bw,2,N,heads]
                w0, b0, w2, b2 = mix
                mo = F.gelu(mo @ w0.t() + b0) @ w2.reshape(-1, 1) + b2                      # [Bw,2,N,1]
                mo = _win_reverse(mo.squeeze(-1).permute(0, 2, 1).reshape(g.rows, 2), g)    # [B2,H,W,2]
                B = g.B2 // 2
                mv = motion.view()
                new = torch.cat([mo[:B], mo[B:]], -1)
                if rows is None:
                    mv[..., motion_off : motion_off + 4] = new
                else:
                    m = _win_token_mask(g, rows)
                    for half in range(2):      # frame-0 queries write channels 0-1, frame-1 queries channels 2-3
                        mh = m[half * B : (half + 1) * B]
                        sl = mv[..., motion_off + 2 * half : motion_off + 2 * half + 2]
                        sl[mh] = new[..., 2 * half : 2 * half + 2][mh]
        self._emit(run)

    def dwconv_gelu(self, x: Map, out: Map, w9c, bias, rows=None):
        def run():
            c = x.C
            wt = w9c.t().reshape(c, 1, 3, 3)
            y = F.conv2d(x.view().permute(0, 3, 1, 2), wt, bias, padding=1, groups=c)
            _put(out.view(), F.gelu(y).permute(0, 2, 3, 1), rows)
        self._emit(run)

    def mlp_tail_ok(self, hidden: Map, fc2: PackedGemm, residual: Map, out: Map) -> bool:
        return fc2.ksize == 1 and list(fc2.split) == [hidden.C] and fc2.Cout % 32 == 0 and hidden.C % 32 == 0

    def mlp_tail(self, hidden: Map, dw_w, dw_b, fc2: PackedGemm, residual: Map, out: Map, rows=None):
        """Contract of atmvfi_mlp_tail = atmvfi_dwconv3x3_gelu followed by the fc2 gemm_conv with the block residual."""
        h2 = self.new_map(hidden.B, hidden.H, hidden.W, hidden.C)
        self.dwconv_gelu(hidden, h2, dw_w, dw_b, rows=rows)
        self.gemm_conv([h2], fc2, out, act=False, residual=residual, rows=rows)

    def flow_warp_nchw(self, img, flow, out, rows=None):
        self._emit(lambda: _put(out, _warp(img, flow), rows, 2))

    def flow_warp_nhwc(self, src: Map, head: Map, flow_off, out: Map, rows=None):
        def run():
            fl = head.view()[..., flow_off : flow_off + 2].permute(0, 3, 1, 2)
            _put(out.view(), _warp(src.view().permute(0, 3, 1, 2), fl).permute(0, 2, 3, 1), rows)
        self._emit(run)

    def warp_blend(self, im0, im1, head: Map, w0, w1, it, flow0=None, flow1=None, occ1=None, occ2=None, rows=None):
        def run():
            hv = head.view().permute(0, 3, 1, 2)
            a, b = _warp(im0, hv[:, 0:2]), _warp(im1, hv[:, 2:4])
            m = torch.sigmoid(hv[:, 4:5])
            _put(w0, a, rows, 2); _put(w1, b, rows, 2); _put(it, m * a + (1 - m) * b, rows, 2)
            if flow0 is not None: _put(flow0, hv[:, 0:2], rows, 2)
            if flow1 is not None: _put(flow1, hv[:, 2:4], rows, 2)
            if occ1 is not None: _put(occ1, m, rows, 2)
            if occ2 is not None: _put(occ2, 1 - m, rows, 2)
        self._emit(run)

    def resize(self, x, out, scale=1.0, rows=None):
        self._emit(lambda: _put(out, F.interpolate(x, size=out.shape[-2:], mode="bilinear", align_corners=True) * scale, rows, 2))

    def nchw_to_nhwc(self, x, out: Map, zero_fill_to=0, rows=None):
        def run():
            c = x.shape[1]
            _put(out.t[..., out.c0 : out.c0 + c], x.permute(0, 2, 3, 1), rows)
            if zero_fill_to > out.c0 + c:
                z = out.t[..., out.c0 + c : zero_fill_to]
                _put(z, torch.zeros_like(z), rows)
        self._emit(run)

    def nhwc_to_nchw(self, x: Map, out):
        self._emit(lambda: out.copy_(x.view().permute(0, 3, 1, 2)))

    def l1_mean(self, a, b, out, scratch):
        self._emit(lambda: out.copy_((a - b).abs().mean(dim=[1, 2, 3]).reshape(out.shape)))

    def select3(self, losses, cands, out):
        def run():
            for i in range(out.shape[0]):
                l = [float(x.reshape(-1)[i]) for x in losses]
                k = 0 if l[0] == min(l) else (1 if l[1] == min(l) else 2)
                out[i] = cands[k][i]
        self._emit(run)

    def copy_map(self, src: Map, dst: Map):
        self._emit(lambda: dst.t.copy_(src.t))

    def residual_finish(self, res: Map, it, it_sum, it_clamped, rows=None):
        def run():
            s = it + (2 * torch.sigmoid(res.view()[..., :3].permute(0, 3, 1, 2)) - 1)
            if it_sum is not None: _put(it_sum, s, rows, 2)
            _put(it_clamped, s.clamp(0, 1), rows, 2)
        self._emit(run)
