"""Operator layer between the host orchestration (engine.py) and the C ABI (include/atmvfi.h).

``Map`` describes a channels-last feature map (or a channel / batch slice of one) by tensor + offsets, so
that every torch.cat / slice / rearrange of the reference becomes pointer arithmetic.  ``CudaOps``
marshals those descriptions into C-ABI calls on the current CUDA stream; with ``recording`` set, calls
are appended to a launch list instead (the engine replays the list, eagerly or under a CUDA graph).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence, Tuple

import torch

from . import _lib


def round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


class Map:
    """NHWC view (fp32, or fp16 in the ATMVFI_F16 mode): tensor ``t`` of shape [B, H, W, pitch]; this view covers channels [c0, c0+C)."""

    __slots__ = ("t", "c0", "C")

    def __init__(self, t: torch.Tensor, c0: int = 0, C: Optional[int] = None):
        assert t.dim() == 4 and t.dtype in (torch.float32, torch.float16)
        assert t.stride(3) == 1 and t.stride(2) == t.shape[3] and t.stride(1) == t.shape[2] * t.shape[3]
        self.t, self.c0 = t, c0
        self.C = t.shape[3] - c0 if C is None else C
        assert 0 <= c0 and c0 + self.C <= t.shape[3]

    B = property(lambda s: s.t.shape[0])
    H = property(lambda s: s.t.shape[1])
    W = property(lambda s: s.t.shape[2])
    pitch = property(lambda s: s.t.shape[3])

    def chan(self, c0: int, C: int) -> "Map":
        assert c0 + C <= self.C
        return Map(self.t, self.c0 + c0, C)

    def batch(self, b0: int, nb: int) -> "Map":
        return Map(self.t[b0 : b0 + nb], self.c0, self.C)

    def rows(self) -> "Map":
        """Same memory seen as [1, 1, B*H*W, pitch] (token rows for the linear layers)."""
        assert self.t.is_contiguous()
        return Map(self.t.view(1, 1, -1, self.t.shape[3]), self.c0, self.C)

    def grid(self, B: int, H: int, W: int) -> "Map":
        assert self.t.is_contiguous() and B * H * W == self.t.shape[0] * self.t.shape[1] * self.t.shape[2]
        return Map(self.t.view(B, H, W, self.t.shape[3]), self.c0, self.C)

    @property
    def ptr(self) -> int:
        return self.t.data_ptr() + self.t.element_size() * self.c0

    @property
    def esize(self) -> int:
        return self.t.element_size()

    @property
    def half(self) -> bool:
        return self.t.dtype == torch.float16

    def view(self) -> torch.Tensor:
        return self.t[..., self.c0 : self.c0 + self.C]

    @property
    def nrows(self) -> int:
        return self.t.shape[0] * self.t.shape[1] * self.t.shape[2]


@dataclass(frozen=True)
class WinGeom:
    B2: int
    H: int
    W: int
    ws: int
    shift: int

    @property
    def Hp(self) -> int:
        return -(-self.H // self.ws) * self.ws

    @property
    def Wp(self) -> int:
        return -(-self.W // self.ws) * self.ws

    @property
    def pad_top(self) -> int:
        return (self.Hp - self.H) // 2

    @property
    def pad_left(self) -> int:
        return (self.Wp - self.W) // 2

    @property
    def rows(self) -> int:
        return self.B2 * self.Hp * self.Wp

    def c(self) -> _lib.WindowGeom:
        return _lib.WindowGeom(self.B2, self.H, self.W, self.ws, self.shift, self.Hp, self.Wp, self.pad_top, self.pad_left)


@dataclass
class PackedGemm:
    """One GEMM-shaped layer after host-side packing (pack.py)."""
    name: str
    ksize: int
    split: Sequence[int]              # channels per source, in order
    Cout: int
    shuffle: bool                     # ConvTranspose2d k2 s2
    w32: torch.Tensor                 # [K, ldw] fp32, K = taps*sum(split)
    bias: Optional[torch.Tensor]
    prelu: Optional[torch.Tensor]
    wtc: Optional[torch.Tensor] = None   # tcgen05 layout, filled by pack.py when the TF32 path is enabled
    wtc3: Optional[torch.Tensor] = None  # 3xTF32 layout (hi | lo chunk pairs)
    wtc16: Optional[torch.Tensor] = None  # fp16 layout (64-channel chunks)
    tc_meta: Optional[dict] = None


def _p(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


Rows = Optional[Tuple[int, int]]      # row window [y0, y1) of every image of the op's output grid; None = all rows


def _yy(rows: Rows) -> Tuple[int, int]:
    if rows is None:
        return 0, 0
    y0, y1 = rows
    assert 0 <= y0 < y1, rows             # an empty window is skipped by the caller, (0, 0) means "all" in the C ABI
    return int(y0), int(y1)


class CudaOps:
    """Launches the sm_100a kernels.  No fallback: construction fails without the library or a CUDA device."""

    def __init__(self, device: torch.device, precision: int = _lib.FP32):
        if device.type != "cuda":
            raise _lib.AtmvfiError(
                f"ATM-VFI B200 kernels need a CUDA device, got '{device}'. There is no CPU fallback; "
                "use the reference implementation for CPU inference.")
        self.lib = _lib.load()
        self.device = device
        self.precision = precision
        self.recording: Optional[List] = None
        self._keep: List = []
        self.launches = 0
        self.round_outputs: Optional[bool] = None
        # optional allocator (shape, zero) -> fp32 tensor; the row-slab mode places every buffer in an IPC-shared arena
        self.allocator: Optional[Callable] = None
        # TF32 mode: the fused q|k|v linear writes the head-major layout that the tcgen05 attention kernel fetches with TMA
        import os
        self.qkv_head_major = precision in (_lib.TF32, _lib.F16) and os.environ.get("ATMVFI_QKV_HEADS", "1") != "0"
        # ATMVFI_F16: channels-last feature maps are stored as fp16 (flows, masks, images, q|k|v and the motion heads stay fp32)
        self.act_f16 = precision == _lib.F16
        self.qkv_head_major_min_hd = int(os.environ.get("ATMVFI_QKV_HEADS_MIN_HD", "0"))      # see engine.transformer_block

    # -- memory -------------------------------------------------------------------------------
    def _alloc(self, shape, zero: bool, dtype=torch.float32) -> torch.Tensor:
        if self.allocator is not None:
            return self.allocator(tuple(shape), zero, dtype) if dtype != torch.float32 else self.allocator(tuple(shape), zero)
        return (torch.zeros if zero else torch.empty)(tuple(shape), device=self.device, dtype=dtype)

    def new_map(self, B: int, H: int, W: int, C: int, zero: bool = False, f32: bool = False) -> Map:
        """``f32``: keep this map fp32 in the fp16-activation mode too (flows, motion, q|k|v, final residual)."""
        half = self.act_f16 and not f32
        pitch = round_up(C, 8 if half else 4)              # rows start on 16-byte boundaries (TMA strides)
        return Map(self._alloc((B, H, W, pitch), zero or pitch != C, torch.float16 if half else torch.float32), 0, C)

    def new_win_map(self, g: "WinGeom", C: int, f32: bool = False) -> Map:
        """Window-major token rows [1, 1, B2*Hp*Wp, C] (what window_partition produces, attention.py:8-15)."""
        return self.new_map(1, 1, g.rows, C, f32=f32)

    def to_act(self, m: Map) -> Map:
        """The activation-typed twin of a small fp32 map that is also a GEMM source (per-token motion, 5-channel motion head):
        the map itself in the fp32-storage modes, an fp16 copy (one cast launch) in the fp16 mode."""
        if not self.act_f16 or m.half:
            return m
        out = self.new_map(m.B, m.H, m.W, m.C)
        self._emit("atmvfi_cast_f32_to_f16", (m.ptr, m.pitch, out.ptr, out.pitch, m.nrows, m.C, out.pitch), keep=(m, out))
        return out

    def new_planar(self, *shape: int) -> torch.Tensor:
        return self._alloc(shape, False)

    def replicated(self):
        """Context in which ops are computed for ALL rows on every rank (a no-op without row slabs)."""
        import contextlib
        return contextlib.nullcontext()

    # -- launch plumbing ----------------------------------------------------------------------
    def set_rounding(self) -> None:
        """tcgen05 kind::tf32 truncates its operands: in that mode producers round feature maps to the nearest TF32 value.
        ``round_outputs`` (None = by precision) lets the per-operator tests look at the un-rounded accumulators."""
        # (fp16 mode: fp32 side outputs - q|k|v for the tf32 attention MMAs - are rounded too; fp16 stores ignore the flag)
        on = self.precision in (_lib.TF32, _lib.F16) if self.round_outputs is None else bool(self.round_outputs)
        self.lib.atmvfi_set_output_rounding(1 if on else 0)
        self.lib.atmvfi_set_activation_f16(1 if self.act_f16 else 0)

    def _emit(self, name: str, args: tuple, keep=()):
        fn = getattr(self.lib, name)
        if self.recording is not None:
            self.recording.append((name, fn, args, keep))
        else:
            self.set_rounding()
            _lib.check(fn(*args, torch.cuda.current_stream(self.device).cuda_stream), name)
            self.launches += 1

    def emit_host(self, fn: Callable[[], None]) -> None:
        """Record (or run) a host-side callable in launch order; used by test transports, never by the product path."""
        if self.recording is not None:
            self.recording.append(("host", lambda *a: (fn(), 0)[1], (), None))
        else:
            fn()

    def replay(self, records, stream: Optional[int] = None) -> None:
        st = torch.cuda.current_stream(self.device).cuda_stream if stream is None else stream
        self.set_rounding()
        for name, fn, args, _ in records:
            rc = fn(*args, st)
            if rc:
                _lib.check(rc, name)
        self.launches += self.count_launches(records)

    @staticmethod
    def count_launches(records) -> int:
        # one kernel per record, except attention-with-motion which also launches the head-mix kernel
        def n(name, args):
            if name == "atmvfi_window_attention":
                return 2 if args[13] is not None else 1
            if name == "atmvfi_window_attention_tc":
                return 2 if args[14] is not None else 1
            if name == "atmvfi_l1_mean":
                return 2
            return 1
        return sum(n(name, args) for name, _, args, _ in records)

    # -- GEMM-shaped layers -------------------------------------------------------------------
    def gemm_conv(self, srcs: Sequence[Map], w: PackedGemm, out: Map, *, stride: int = 1, dil: int = 1,
                  act: bool = True, residual: Optional[Map] = None, out2: Optional[Map] = None,
                  prelu2: Optional[torch.Tensor] = None, win: Optional[WinGeom] = None,
                  precision: Optional[int] = None, rows: Rows = None, qkv_heads: int = 0,
                  out_f32: bool = False, head32: Optional[Map] = None, head32_c0: int = 0):
        assert [s.C for s in srcs] == list(w.split), (w.name, [s.C for s in srcs], w.split)
        s0 = srcs[0]
        for s in srcs:
            assert (s.B, s.H, s.W) == (s0.B, s0.H, s0.W), w.name
        pad = dil * (w.ksize - 1) // 2
        Hout = (s0.H + 2 * pad - dil * (w.ksize - 1) - 1) // stride + 1
        Wout = (s0.W + 2 * pad - dil * (w.ksize - 1) - 1) // stride + 1
        d = _lib.GemmConvDesc()
        d.nsrc = len(srcs)
        for i, s in enumerate(srcs):
            d.src[i] = _lib.Src(s.ptr, s.C, s.pitch)
        d.B, d.Hin, d.Win = s0.B, s0.H, s0.W
        d.ksize, d.stride, d.dil = w.ksize, stride, dil
        d.Hout, d.Wout, d.Cout = Hout, Wout, w.Cout
        d.weight, d.ldw = w.w32.data_ptr(), w.w32.shape[1]
        d.bias = _p(w.bias)
        d.prelu = _p(w.prelu) if act else None
        pad4 = None
        tc_prec = (self.precision if precision is None else precision) in (_lib.TF32, _lib.TF32X3, _lib.F16)
        if tc_prec and win is None and not qkv_heads:
            # Tensor-core epilogues read bias / slopes in whole vectors: hand them over padded to a multiple of 32 floats (cached per
            # layer).  Odd channel counts (101, 197, 389, 5, 3 ...) on a map that owns its pad lanes (c0 == 0) may additionally be
            # stored as whole 4-channel vectors (pad_stores) when the layer takes the non-TMA vector epilogue.
            pad4 = self._padded(w, prelu2)
            d.bias, d.prelu, d.param_pad = _p(pad4[0]), (_p(pad4[1]) if act else None), 32
            if w.Cout % 4 and not w.shuffle and out.c0 == 0 and (out2 is None or out2.c0 == 0):
                d.pad_stores = 1
        if residual is not None:
            assert residual.C == w.Cout
            d.residual, d.res_pitch = residual.ptr, residual.pitch
        d.out, d.out_pitch = out.ptr, out.pitch
        assert out.C == w.Cout, (w.name, out.C, w.Cout)
        if out2 is not None:
            d.out2, d.prelu2, d.out2_pitch = out2.ptr, (prelu2 if pad4 is None else pad4[2]).data_ptr(), out2.pitch
        if win is not None:
            d.out_mode, d.win = _lib.OUT_WINDOW_REV, win.c()
            assert s0.nrows == win.rows and out.nrows == win.B2 * win.H * win.W
        elif w.shuffle:
            d.out_mode = _lib.OUT_SHUFFLE2
            assert (out.B, out.H, out.W) == (s0.B, 2 * Hout, 2 * Wout), w.name
        elif qkv_heads:
            # fused q|k|v linear written head-major (include/atmvfi.h ATMVFI_OUT_QKV_HEADS); `out` is the [rows][3C] buffer reinterpreted
            d.out_mode, d.qkv_heads = _lib.OUT_QKV_HEADS, qkv_heads
            assert out.c0 == 0 and out.pitch == w.Cout and w.Cout % (3 * qkv_heads) == 0 and out.nrows == s0.B * Hout * Wout
            assert residual is None and out2 is None and w.ksize == 1 and stride == 1
        else:
            d.out_mode = _lib.OUT_PIXEL
            assert out.nrows == s0.B * Hout * Wout, (w.name, out.t.shape, (s0.B, Hout, Wout))
        d.row_begin, d.row_end = _yy(rows)
        assert d.row_end <= Hout
        prec = self.precision if precision is None else precision
        if prec in (_lib.TF32, _lib.TF32X3) and not self._tc_eligible(srcs, w, out, out2):
            prec = _lib.FP32
        if prec == _lib.F16:
            # fp16 storage: there is no CUDA-core fallback for fp16 maps, every GEMM-shaped layer must fit the tensor-core kernel
            if not self._tc_eligible(srcs, w, out, out2):
                raise _lib.AtmvfiError(f"gemm_conv({w.name}): layer does not fit the tcgen05 kernel (>= 16 input channels, 16-byte aligned operands) "
                                       "and fp16 feature maps have no CUDA-core path")
            assert all(s.half for s in srcs) and (residual is None or residual.half) and (out2 is None or out2.half), w.name
            assert out.half != bool(out_f32), (w.name, "out dtype does not match out_f32")
            d.out_f32 = int(bool(out_f32))
            if head32 is not None:
                assert not head32.half and head32.nrows == out.nrows and head32.C == w.Cout - head32_c0 and win is None and not w.shuffle
                d.head32, d.head32_pitch, d.head32_c0 = head32.ptr, head32.pitch, head32_c0
        else:
            assert not any(s.half for s in srcs) and not out.half and head32 is None, w.name
        d.precision = prec
        planbuf = None
        if prec in (_lib.TF32, _lib.TF32X3, _lib.F16):
            if prec == _lib.F16:
                if w.wtc16 is None:
                    from .pack import pack_tc_f16
                    w.wtc16 = pack_tc_f16(w)
                wt = w.wtc16
            elif prec == _lib.TF32:
                if w.wtc is None:
                    from .pack import pack_tc
                    w.wtc = pack_tc(w)
                wt = w.wtc
            else:
                if w.wtc3 is None:
                    from .pack import pack_tc_x3
                    w.wtc3 = pack_tc_x3(w)
                wt = w.wtc3
            d.weight, d.ldw = wt.data_ptr(), wt.shape[1]
            nbytes = self.lib.atmvfi_gemm_conv_plan_bytes()
            planbuf = C.create_string_buffer(nbytes + 64)
            addr = (C.addressof(planbuf) + 63) & ~63
            _lib.check(self.lib.atmvfi_gemm_conv_plan(C.byref(d), addr), f"gemm_conv_plan({w.name})")
            d.tma_host = addr
        self._emit("atmvfi_gemm_conv", (C.byref(d),), keep=(d, srcs, w, out, residual, out2, prelu2, planbuf, head32, pad4))

    @staticmethod
    def _padded(w: PackedGemm, prelu2: Optional[torch.Tensor]):
        """(bias, prelu, prelu2) of a layer padded to round_up(Cout, 32) + 32 floats (zeros / ones), cached on the packed layer."""
        cache = w.__dict__.setdefault("_pad4", {})
        key = None if prelu2 is None else prelu2.data_ptr()
        if key not in cache:
            n4 = round_up(w.Cout, 32) + 32

            def pad(t, fill):
                if t is None:
                    return None
                o = torch.full((n4,), fill, dtype=torch.float32, device=t.device)
                o[: w.Cout] = t.reshape(-1)[: w.Cout]
                return o
            cache[key] = (pad(w.bias, 0.0), pad(w.prelu, 1.0), pad(prelu2, 1.0))
        return cache[key]

    @staticmethod
    def _tc_eligible(srcs, w, out, out2) -> bool:
        """Layers the tcgen05 kernel takes: 16-byte aligned operands and enough input channels to fill a K chunk."""
        if sum(w.split) < 16:
            return False
        oa = lambda m: 8 if m.half else 16             # fp16 outputs are written as 8-byte vectors
        if any(s.ptr % 16 for s in srcs) or out.ptr % oa(out) or (out2 is not None and out2.ptr % oa(out2)):
            return False
        return True

    def conv3x3_first(self, img: torch.Tensor, w: PackedGemm, out: Map, rows: Rows = None):
        """feat_extracts.0.0 on a planar frame (3 channels) -> NHWC."""
        b, c, h, wd = img.shape
        assert c == 3 and w.ksize == 3 and list(w.split) == [3] and (out.B, out.H, out.W, out.C) == (b, h, wd, w.Cout) and out.c0 == 0
        self._emit("atmvfi_conv3x3_first", (img.data_ptr(), w.w32.data_ptr(), w.w32.shape[1], _p(w.bias), _p(w.prelu), out.ptr, out.pitch, b, h, wd, w.Cout) + _yy(rows),
                   keep=(img, w, out))

    def pack5_planar(self, imgs: Sequence[torch.Tensor], out: Map, rows: Rows = None):
        b, _, h, wd = imgs[0].shape
        assert len(imgs) == 5 and all(t.shape == (b, 3, h, wd) and t.is_contiguous() for t in imgs) and out.c0 == 0 and out.pitch >= 16
        self._emit("atmvfi_pack5_planar", tuple(t.data_ptr() for t in imgs) + (out.ptr, out.pitch, b, h, wd) + _yy(rows), keep=(imgs, out))

    # -- transformer pieces -------------------------------------------------------------------
    def layernorm(self, x: Map, out: Map, gamma: torch.Tensor, beta: torch.Tensor, rows: Rows = None):
        assert x.C == out.C == gamma.numel()
        if rows is None:
            self._emit("atmvfi_layernorm", (x.ptr, x.pitch, out.ptr, out.pitch, x.nrows, x.C, gamma.data_ptr(), beta.data_ptr(), 1e-5),
                       keep=(x, out, gamma, beta))
            return
        # per-token op: a row window is a contiguous run of tokens in every image
        assert (x.B, x.H, x.W) == (out.B, out.H, out.W)
        y0, y1 = _yy(rows)
        for b in range(x.B):
            first = (b * x.H + y0) * x.W
            self._emit("atmvfi_layernorm", (x.ptr + 4 * first * x.pitch, x.pitch, out.ptr + 4 * first * out.pitch, out.pitch, (y1 - y0) * x.W, x.C,
                                            gamma.data_ptr(), beta.data_ptr(), 1e-5), keep=(x, out, gamma, beta))

    def window_gather_ln(self, tok: Map, win: Map, g: WinGeom, gamma: torch.Tensor, beta: torch.Tensor, rows: Rows = None):
        assert tok.nrows == g.B2 * g.H * g.W and win.nrows == g.rows and tok.C == win.C
        gc = g.c()
        self._emit("atmvfi_window_gather_ln", (tok.ptr, tok.pitch, win.ptr, win.pitch, tok.C, C.byref(gc), gamma.data_ptr(), beta.data_ptr(), 1e-5) + _yy(rows),
                   keep=(tok, win, gc, gamma, beta))

    def window_attention(self, qkv: Map, out: Map, g: WinGeom, heads: int, cross: bool, rc: Optional[torch.Tensor] = None,
                         mix: Optional[Sequence[torch.Tensor]] = None, motion: Optional[Map] = None, motion_off: int = 0,
                         scratch: Optional[torch.Tensor] = None, rc_closed_form: bool = False, rows: Rows = None, head_major: bool = False):
        assert qkv.C == 3 * out.C and qkv.nrows == g.rows == out.nrows
        assert not head_major or (qkv.c0 == 0 and qkv.pitch == qkv.C)
        gc = g.c()
        m = [None] * 4 if mix is None else [t.data_ptr() for t in mix]
        tail = (m[0], m[1], m[2], m[3], None if motion is None else motion.ptr, 0 if motion is None else motion.pitch, motion_off, _p(scratch)) + _yy(rows) + (int(head_major),)
        head = (qkv.ptr, qkv.pitch, out.ptr, out.pitch, out.C, heads, C.byref(gc), int(cross), _p(rc))
        keep = (qkv, out, gc, rc, mix, motion, scratch)
        if self.precision in (_lib.TF32, _lib.F16):
            assert not qkv.half and out.half == self.act_f16 and (motion is None or not motion.half)
            self._emit("atmvfi_window_attention_tc", head + (int(rc_closed_form),) + tail, keep=keep)
        else:
            self._emit("atmvfi_window_attention", head + tail, keep=keep)

    def dwconv_gelu(self, x: Map, out: Map, w9c: torch.Tensor, bias: torch.Tensor, rows: Rows = None):
        assert x.c0 == 0 and out.c0 == 0 and x.pitch == out.pitch and x.C == out.C
        self._emit("atmvfi_dwconv3x3_gelu", (x.ptr, out.ptr, x.B, x.H, x.W, x.C, x.pitch, w9c.data_ptr(), bias.data_ptr()) + _yy(rows),
                   keep=(x, out, w9c, bias))

    def mlp_tail_ok(self, hidden: Map, fc2: PackedGemm, residual: Map, out: Map) -> bool:
        """Shapes the fused Mlp tail (atmvfi_mlp_tail) takes; everything else runs dwconv_gelu + gemm_conv."""
        import os
        if self.precision not in (_lib.TF32, _lib.F16) or os.environ.get("ATMVFI_MLP_TAIL", "1") == "0":
            return False
        chunk = 64 if self.act_f16 else 32
        maps = (hidden, residual, out)
        return (hidden.C % chunk == 0 and fc2.Cout % 32 == 0 and fc2.ksize == 1 and list(fc2.split) == [hidden.C]
                and all(m.c0 == 0 and (m.pitch * m.esize) % 16 == 0 and m.half == self.act_f16 for m in maps))

    def mlp_tail(self, hidden: Map, dw_w: torch.Tensor, dw_b: torch.Tensor, fc2: PackedGemm, residual: Map, out: Map, rows: Rows = None):
        """out = residual + fc2(GELU(DWConv3x3(hidden) + dw_b)) + b_fc2 in one launch (include/atmvfi.h atmvfi_mlp_tail)."""
        assert self.mlp_tail_ok(hidden, fc2, residual, out) and all(m.ptr % 16 == 0 for m in (hidden, residual, out))
        assert (hidden.B, hidden.H, hidden.W) == (residual.B, residual.H, residual.W) == (out.B, out.H, out.W) and residual.C == out.C == fc2.Cout
        cache = fc2.__dict__.setdefault("_mlp_tail", {})
        key = (dw_w.data_ptr(), dw_b.data_ptr())
        if key not in cache:
            w10 = torch.cat([dw_w.reshape(9, -1), dw_b.reshape(1, -1)], 0).contiguous().float()
            nb = round_up(fc2.Cout, 32) + 384
            bias = torch.zeros(nb, dtype=torch.float32, device=w10.device)
            if fc2.bias is not None:
                bias[: fc2.Cout] = fc2.bias.reshape(-1)
            cache[key] = (w10, bias)
        w10, bias = cache[key]
        if self.act_f16:
            if fc2.wtc16 is None:
                from .pack import pack_tc_f16
                fc2.wtc16 = pack_tc_f16(fc2)
            wt = fc2.wtc16
        else:
            if fc2.wtc is None:
                from .pack import pack_tc
                fc2.wtc = pack_tc(fc2)
            wt = fc2.wtc
        assert wt.shape[1] == hidden.C and wt.shape[0] >= fc2.Cout
        self._emit("atmvfi_mlp_tail", (hidden.ptr, hidden.pitch, hidden.B, hidden.H, hidden.W, hidden.C, w10.data_ptr(), wt.data_ptr(), wt.shape[0],
                                       bias.data_ptr(), residual.ptr, residual.pitch, out.ptr, out.pitch, fc2.Cout, int(self.precision)) + _yy(rows),
                   keep=(hidden, w10, bias, wt, fc2, residual, out))

    # -- warps, resampling, layout ------------------------------------------------------------
    def flow_warp_nchw(self, img: torch.Tensor, flow: torch.Tensor, out: torch.Tensor, rows: Rows = None):
        b, c, h, w = img.shape
        assert flow.shape == (b, 2, h, w) and img.is_contiguous() and flow.is_contiguous() and out.is_contiguous()
        self._emit("atmvfi_flow_warp_nchw", (img.data_ptr(), flow.data_ptr(), out.data_ptr(), b, c, h, w) + _yy(rows), keep=(img, flow, out))

    def pyramid_warp(self, im0: torch.Tensor, im1: torch.Tensor, flow0: torch.Tensor, flow1: torch.Tensor, upsample: bool,
                     out0: torch.Tensor, out1: torch.Tensor, flow0_out: Optional[torch.Tensor] = None, flow1_out: Optional[torch.Tensor] = None,
                     rows: Rows = None):
        """One level of the global-motion pyramid warp for both frames (include/atmvfi.h atmvfi_pyramid_warp): optional x2 flow
        up-sampling fused into the two backward warps."""
        b, c, h, w = im0.shape
        fs = (b, 2, h // 2, w // 2) if upsample else (b, 2, h, w)
        assert c == 3 and im1.shape == im0.shape == out0.shape == out1.shape and tuple(flow0.shape) == fs == tuple(flow1.shape)
        assert all(t is None or (tuple(t.shape) == (b, 2, h, w) and t.is_contiguous()) for t in (flow0_out, flow1_out))
        self._emit("atmvfi_pyramid_warp", (im0.data_ptr(), im1.data_ptr(), flow0.data_ptr(), flow1.data_ptr(), int(bool(upsample)), out0.data_ptr(),
                                           out1.data_ptr(), _p(flow0_out), _p(flow1_out), b, h, w) + _yy(rows),
                   keep=(im0, im1, flow0, flow1, out0, out1, flow0_out, flow1_out))

    @staticmethod
    def _row_owners(owners, H: int):
        """[(row_lo, row_hi, byte_delta)] covering [0, H) -> the C struct (rows held by other GPUs, read in place over NVLink)."""
        ro = _lib.RowOwners()
        assert 0 < len(owners) <= 2 * _lib.P2P_MAX_PEERS and owners[0][0] == 0 and owners[-1][1] == H
        ro.nseg = len(owners)
        for i, (lo, hi, delta) in enumerate(owners):
            assert i == 0 or owners[i - 1][1] == lo
            ro.row_lo[i], ro.byte_delta[i] = lo, delta
        ro.row_lo[len(owners)] = H
        return ro

    def flow_warp_nhwc(self, src: Map, head: Map, flow_off: int, out: Map, rows: Rows = None, owners=None):
        """owners: [(row_lo, row_hi, byte_delta)] - source rows held by other GPUs, read in place over NVLink (row slabs)."""
        assert (src.B, src.H, src.W) == (head.B, head.H, head.W) == (out.B, out.H, out.W) and src.C == out.C
        args = (src.ptr, src.pitch, head.ptr, head.pitch, flow_off, out.ptr, out.pitch, src.B, src.C, src.H, src.W) + _yy(rows)
        if owners is None:
            self._emit("atmvfi_flow_warp_nhwc", args, keep=(src, head, out))
            return
        ro = self._row_owners(owners, src.H)
        self._emit("atmvfi_flow_warp_nhwc_p2p", args + (C.byref(ro),), keep=(src, head, out, ro))

    def warp_blend(self, im0, im1, head: Map, w0, w1, it, flow0=None, flow1=None, occ1=None, occ2=None, rows: Rows = None, owners=None):
        b, _, h, w = im0.shape
        assert (head.B, head.H, head.W) == (b, h, w) and head.C >= 5
        args = (im0.data_ptr(), im1.data_ptr(), head.ptr, head.pitch, 0, w0.data_ptr(), w1.data_ptr(), it.data_ptr(),
                _p(flow0), _p(flow1), _p(occ1), _p(occ2), b, h, w) + _yy(rows)
        keep = (im0, im1, head, w0, w1, it, flow0, flow1, occ1, occ2)
        if owners is None:
            self._emit("atmvfi_warp_blend", args, keep=keep)
        else:
            ro = self._row_owners(owners, h)
            self._emit("atmvfi_warp_blend_p2p", args + (C.byref(ro),), keep=keep + (ro,))

    def resize(self, x: torch.Tensor, out: torch.Tensor, scale: float = 1.0, rows: Rows = None):
        assert x.is_contiguous() and out.is_contiguous() and x.shape[:2] == out.shape[:2]
        self._emit("atmvfi_resize_bilinear_ac", (x.data_ptr(), out.data_ptr(), x.shape[0] * x.shape[1], x.shape[2], x.shape[3], out.shape[2], out.shape[3], float(scale)) + _yy(rows),
                   keep=(x, out))

    def nchw_to_nhwc(self, x: torch.Tensor, out: Map, zero_fill_to: int = 0, rows: Rows = None):
        b, c, h, w = x.shape
        assert (out.B, out.H, out.W) == (b, h, w) and out.C >= c and x.is_contiguous()
        self._emit("atmvfi_nchw_to_nhwc", (x.data_ptr(), out.t.data_ptr(), out.pitch, out.c0, b, c, h, w, zero_fill_to) + _yy(rows), keep=(x, out))

    def nhwc_to_nchw(self, x: Map, out: torch.Tensor):
        b, c, h, w = out.shape
        assert (x.B, x.H, x.W, x.C) == (b, h, w, c) and out.is_contiguous()
        self._emit("atmvfi_nhwc_to_nchw", (x.ptr, x.pitch, out.data_ptr(), b, c, h, w), keep=(x, out))

    # -- multi-scale global-motion ensemble (network_base.py:548-615) ---------------------------------
    def l1_mean(self, a: torch.Tensor, b: torch.Tensor, out: torch.Tensor, scratch: torch.Tensor):
        n = a[0].numel()
        assert a.shape == b.shape and out.numel() == a.shape[0] and scratch.numel() >= self.lib.atmvfi_l1_mean_scratch_floats(a.shape[0])
        self._emit("atmvfi_l1_mean", (a.data_ptr(), b.data_ptr(), out.data_ptr(), scratch.data_ptr(), a.shape[0], n), keep=(a, b, out, scratch))

    def select3(self, losses: Sequence[torch.Tensor], cands: Sequence[torch.Tensor], out: torch.Tensor):
        assert len(losses) == len(cands) == 3 and all(c.shape == out.shape and c.is_contiguous() for c in cands)
        self._emit("atmvfi_select_min3", tuple(l.data_ptr() for l in losses) + tuple(c.data_ptr() for c in cands) +
                   (out.data_ptr(), out.shape[0], out[0].numel()), keep=(losses, cands, out))

    def copy_map(self, src: Map, dst: Map):
        """dst <- src for two whole (contiguous) maps of equal shape."""
        assert src.t.shape == dst.t.shape and src.t.is_contiguous() and dst.t.is_contiguous() and src.c0 == dst.c0 == 0
        self._emit("atmvfi_copy", (dst.t.data_ptr(), src.t.data_ptr(), src.t.numel() * src.t.element_size()), keep=(src, dst))

    def residual_finish(self, res: Map, it, it_sum, it_clamped, rows: Rows = None):
        b, _, h, w = it.shape
        self._emit("atmvfi_residual_finish", (res.ptr, res.pitch, it.data_ptr(), _p(it_sum), it_clamped.data_ptr(), b, h, w) + _yy(rows),
                   keep=(res, it, it_sum, it_clamped))

    def u8_to_planar(self, src_u8: torch.Tensor, out: torch.Tensor, H, W, Hp, Wp, top, left, bgr: bool):
        self._emit("atmvfi_u8_to_planar", (src_u8.data_ptr(), out.data_ptr(), H, W, Hp, Wp, top, left, int(bgr)), keep=(src_u8, out))

    def planar_to_u8(self, src: torch.Tensor, out_u8: torch.Tensor, H, W, Hp, Wp, top, left, bgr: bool):
        self._emit("atmvfi_planar_to_u8", (src.data_ptr(), out_u8.data_ptr(), H, W, Hp, Wp, top, left, int(bgr)), keep=(src, out_u8))
