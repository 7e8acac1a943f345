"""Spatial row-slab execution of ONE frame pair on several GPUs (BASELINE.json configs[3]).

``SlabOps`` wraps an operator backend (``ops.CudaOps``) behind the same interface that ``engine.Plan`` records
against.  Every rank builds the SAME plan with full-size buffers, but launches each operator only on its own row
window (``rows=`` of the C ABI: padding, align_corners scales, window geometry and warp coordinates stay global).
While the plan is built, the wrapper tracks for EVERY rank which rows of every buffer are valid there and, in
front of each consumer, schedules the neighbour rows that are missing as an exchange site: the rank that holds
the rows pushes them into the consumer's copy of the buffer (``transport.exchange``: NVLink P2P stores + flag,
csrc/p2p.cu).  All ranks derive the same schedule, so no negotiation happens at run time.

Row dependencies by operator class:
  * per-token / per-pixel ops                    - the same rows;
  * 3x3 convs (stride 1/2/4, dilation), DWConv    - a fixed halo;  k2s2 transposed conv - none;
  * align_corners resizes                         - the source rows the bilinear taps touch;
  * window attention                              - a row of windows is owned by the rank that owns its first real
    token row; under the cyclic shift (and the centre padding of the 1/16 grid) its tokens straddle the slab
    boundary, so input rows are pulled in and the rows produced for a neighbour are pushed back;
  * backward warps                                - data-dependent reach: the 3-channel image pyramids are kept
    whole on every rank (``replicated()``), the 1/8-resolution feature maps are all-gathered in front of the warp.
"""
from __future__ import annotations

import bisect
import contextlib
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .ops import Map, PackedGemm, WinGeom

Interval = Tuple[int, int]
RowSet = List[Interval]          # sorted, disjoint, non-empty intervals


# ------------------------------------------------------------------------------------------------ interval sets
def rs_norm(iv: Sequence[Interval]) -> RowSet:
    out: RowSet = []
    for lo, hi in sorted((a, b) for a, b in iv if b > a):
        if out and lo <= out[-1][1]:
            out[-1] = (out[-1][0], max(out[-1][1], hi))
        else:
            out.append((lo, hi))
    return out


def rs_union(a: RowSet, b: RowSet) -> RowSet:
    return rs_norm(list(a) + list(b))


def rs_sub(a: RowSet, b: RowSet) -> RowSet:
    out: RowSet = []
    for lo, hi in a:
        cur = lo
        for blo, bhi in b:
            if bhi <= cur or blo >= hi:
                continue
            if blo > cur:
                out.append((cur, blo))
            cur = max(cur, bhi)
            if cur >= hi:
                break
        if cur < hi:
            out.append((cur, hi))
    return out


def rs_and(a: RowSet, b: RowSet) -> RowSet:
    out: RowSet = []
    for lo, hi in a:
        for blo, bhi in b:
            l, h = max(lo, blo), min(hi, bhi)
            if h > l:
                out.append((l, h))
    return rs_norm(out)


# ------------------------------------------------------------------------------------------------ window geometry
def win_row_tokens(g: WinGeom, k: int) -> List[int]:
    """Real token rows (un-padded grid) seen by window row k, in window order (attention.py:273-316)."""
    rows = []
    for i in range(k * g.ws, (k + 1) * g.ws):
        y = (i + g.shift) % g.Hp - g.pad_top
        if 0 <= y < g.H:
            rows.append(y)
    return rows


def win_rows_to_tokens(g: WinGeom, k0: int, k1: int) -> RowSet:
    return rs_norm([(y, y + 1) for k in range(k0, k1) for y in win_row_tokens(g, k)])


# ------------------------------------------------------------------------------------------------ buffers
class BufInfo:
    """One allocation, identical on every rank.  ``rows3`` views it as [image*plane, rows, row elements]."""

    def __init__(self, idx: int, kind: str, t: torch.Tensor, images: int, planes: int, H: int, g: Optional[WinGeom], world: int):
        self.idx, self.kind, self.t, self.images, self.planes, self.H, self.g = idx, kind, t, images, planes, H, g
        self.rows3 = t.view(images * planes, H, -1)
        self.row_bytes = self.rows3.shape[2] * 4
        self.image_bytes = planes * H * self.row_bytes
        # have[rank][image]: rows valid on that rank.  A buffer that no operator has written yet is an external input
        # (the frames): every rank holds all of it.
        self.external = True
        self.have: List[List[RowSet]] = [[[] for _ in range(images)] for _ in range(world)]
        # order key (SlabOps._keys) of the last launch on THIS rank after which rows of the buffer became valid here: a push of
        # such rows may be issued right behind that launch
        self.ready_key = 0.0

    def grid_map(self, like: Map) -> Map:
        """``like`` (any reshaped view of this buffer, e.g. ``Map.rows()``) on the buffer's native [B,H,W,pitch] grid."""
        assert self.kind in ("nhwc", "win")
        off = like.t.data_ptr() - self.t.data_ptr()
        assert off % self.image_bytes == 0
        b0 = off // self.image_bytes
        nb = like.t.numel() * 4 // self.image_bytes
        assert nb * self.image_bytes == like.t.numel() * 4
        t4 = self.t4[b0 : b0 + nb]
        return Map(t4, like.c0, like.C)


class Push:
    __slots__ = ("buf", "img0", "nimg", "lo", "hi", "src", "dst")

    def __init__(self, buf: BufInfo, img0: int, nimg: int, lo: int, hi: int, src: int, dst: int):
        self.buf, self.img0, self.nimg, self.lo, self.hi, self.src, self.dst = buf, img0, nimg, lo, hi, src, dst

    def nbytes(self) -> int:
        return self.nimg * self.buf.planes * (self.hi - self.lo) * self.buf.row_bytes


def slab_bounds(H: int, div: int, world: int, align: int = 64) -> List[int]:
    """Row boundaries of the slabs in full-resolution pixels: multiples of ``align`` (whole 8x8 windows of the 1/8 grid)
    when the image has at least one such block per rank, else multiples of ``div`` (the coarsest grid's pixel pitch)."""
    unit = align if (H % align == 0 and H // align >= world) else div
    n = H // unit
    if n < world:
        raise RuntimeError(f"row slabs: an image of {H} rows has only {n} blocks of {unit} rows for {world} ranks")
    return [unit * (n * r // world) for r in range(world + 1)]


class SlabOps:
    """Same interface as ``ops.CudaOps``; see the module docstring."""

    def __init__(self, backend, rank: int, world: int, transport, gather: str = "all"):
        assert 0 <= rank < world and gather in ("all", "I_t", "none")
        self.backend, self.rank, self.world, self.transport, self.gather = backend, rank, world, transport, gather
        self.bufs: List[BufInfo] = []
        self._starts: List[int] = []          # sorted data_ptr of every buffer (bisect lookup)
        self._by_start: Dict[int, BufInfo] = {}
        self._replicated = 0
        self.bounds: Optional[List[int]] = None
        self.H = 0
        self.site = 0
        self.stats = {"sites": 0, "pushed_bytes": 0, "received_bytes": 0}
        self._keys: List[float] = []
        self._next_key = 0.0
        transport.attach(self)

    # ---- plumbing shared with the backend -------------------------------------------------------
    def _set_recording(self, v):
        self.backend.recording = v
        self._keys, self._next_key = [], 0.0          # order keys of the records of the list being built (see _run)

    recording = property(lambda s: s.backend.recording, _set_recording)

    def _sync_keys(self, key: Optional[float] = None) -> float:
        """Give every record appended since the last call its order key: the next integer, or ``key`` (a push that is to be
        issued right behind the launch that produced its rows).  Returns the key of the newest record."""
        rec = self.backend.recording
        if rec is None:
            return 0.0
        while len(self._keys) < len(rec):
            if key is None:
                self._next_key += 1.0
                self._keys.append(self._next_key)
            else:
                self._keys.append(key)
        return self._keys[-1] if self._keys else 0.0

    def _reorder(self) -> None:
        """Move every early push to its place (stable sort by order key); called once, at the end of a plan build."""
        rec = self.backend.recording
        if rec is None or not getattr(self.transport, "split_sites", False):
            return
        self._sync_keys()
        order = sorted(range(len(rec)), key=lambda i: self._keys[i])
        rec[:] = [rec[i] for i in order]
        self._keys = [self._keys[i] for i in order]
    launches = property(lambda s: s.backend.launches)
    lib = property(lambda s: s.backend.lib)
    device = property(lambda s: s.backend.device)
    precision = property(lambda s: s.backend.precision, lambda s, v: setattr(s.backend, "precision", v))
    qkv_head_major = property(lambda s: bool(getattr(s.backend, "qkv_head_major", False)))
    qkv_head_major_min_hd = property(lambda s: getattr(s.backend, "qkv_head_major_min_hd", 48))

    def replay(self, records, stream=None):
        return self.backend.replay(records, stream)

    def set_rounding(self):
        return self.backend.set_rounding()

    def count_launches(self, records):
        return self.backend.count_launches(records)

    def begin_plan(self, B: int, H: int, W: int, glob: bool) -> None:
        self.H = H
        import os
        self.bounds = slab_bounds(H, 16 if glob else 8, self.world, int(os.environ.get("ATMVFI_SLAB_ALIGN", "16")))
        self.transport.step_begin()

    @contextlib.contextmanager
    def replicated(self):
        self._replicated += 1
        try:
            yield
        finally:
            self._replicated -= 1

    # ---- allocation -------------------------------------------------------------------------------
    def _register(self, kind: str, t: torch.Tensor, images: int, planes: int, H: int, g: Optional[WinGeom] = None) -> BufInfo:
        b = BufInfo(len(self.bufs), kind, t, images, planes, H, g, self.world)
        if kind == "nhwc":
            b.t4 = t
        elif kind == "win":
            b.t4 = t.view(g.B2, g.Hp // g.ws, g.ws * g.Wp, t.shape[-1])
        self.bufs.append(b)
        bisect.insort(self._starts, t.data_ptr())
        self._by_start[t.data_ptr()] = b
        return b

    act_f16 = False                # row slabs run on fp32 feature maps (tf32 / fp32 / fp32x3)

    def to_act(self, m: Map) -> Map:
        return m

    def new_map(self, B, H, W, C, zero=False, f32=False) -> Map:
        m = self.backend.new_map(B, H, W, C, zero)
        self._register("nhwc", m.t, B, 1, H)
        return m

    def new_win_map(self, g: WinGeom, C, f32=False) -> Map:
        m = self.backend.new_win_map(g, C)
        self._register("win", m.t, g.B2, 1, g.Hp // g.ws, g)
        return m

    def new_planar(self, *shape) -> torch.Tensor:
        t = self.backend.new_planar(*shape)
        b, c, h, w = shape
        self._register("planar", t, b, c, h)
        return t

    def _buf(self, x) -> Tuple[BufInfo, int, int]:
        """(buffer, first image, image count) of a Map or planar tensor view."""
        t = x.t if isinstance(x, Map) else x
        p = t.data_ptr()
        i = bisect.bisect_right(self._starts, p) - 1
        assert i >= 0, "tensor was not allocated through SlabOps"
        b = self._by_start[self._starts[i]]
        off = p - b.t.data_ptr()
        assert 0 <= off < b.images * b.image_bytes and off % b.image_bytes == 0, "view does not start on an image boundary"
        nimg = t.numel() * 4 // b.image_bytes
        assert nimg * b.image_bytes == t.numel() * 4 and nimg >= 1
        return b, off // b.image_bytes, nimg

    # ---- ownership --------------------------------------------------------------------------------
    def own(self, b: BufInfo, r: int) -> Interval:
        """Rows of ``b`` that rank r produces (for a window-major buffer: rows of windows)."""
        if self._replicated:
            return (0, b.H)
        if b.kind == "win":
            return self._win_own(b.g, r)
        scale = self.H // b.H
        assert scale * b.H == self.H and self.bounds[r] % scale == 0 and self.bounds[r + 1] % scale == 0, (self.H, b.H)
        return (self.bounds[r] // scale, self.bounds[r + 1] // scale)

    def _win_own(self, g: WinGeom, r: int) -> Interval:
        scale = self.H // g.H
        a, b = self.bounds[r] // scale, self.bounds[r + 1] // scale
        ks = [k for k in range(g.Hp // g.ws) if a <= win_row_tokens(g, k)[0] < b]
        if not ks:
            return (0, 0)
        assert ks == list(range(ks[0], ks[-1] + 1))
        return (ks[0], ks[-1] + 1)

    # ---- the scheduler ------------------------------------------------------------------------------
    def _run(self, reads, writes, launch) -> None:
        """reads / writes: per rank, list of (view, RowSet).  Schedules the exchange, launches my part, updates validity."""
        pushes: List[Push] = []
        for p in range(self.world):
            for view, need in reads[p]:
                if not need:
                    continue
                b, i0, ni = self._buf(view)
                if b.external:
                    continue                           # an input of the plan: valid everywhere
                for img in range(i0, i0 + ni):
                    missing = rs_sub(need, b.have[p][img])
                    if not missing:
                        continue
                    # nearest ranks first: neighbours hold halo rows
                    for q in sorted((q for q in range(self.world) if q != p), key=lambda q: (abs(q - p), q)):
                        got = rs_and(missing, b.have[q][img])
                        for lo, hi in got:
                            pushes.append(Push(b, img, 1, lo, hi, q, p))
                        missing = rs_sub(missing, got)
                        if not missing:
                            break
                    assert not missing, f"rows {missing} of buffer {b.idx} (image {img}) exist on no rank"
        if pushes:
            pushes = self._merge(pushes)
            for ps in pushes:                          # validity after the exchange
                for img in range(ps.img0, ps.img0 + ps.nimg):
                    ps.buf.have[ps.dst][img] = rs_union(ps.buf.have[ps.dst][img], [(ps.lo, ps.hi)])
            mine_out = [ps for ps in pushes if ps.src == self.rank]
            mine_in = [ps for ps in pushes if ps.dst == self.rank]
            self.stats["sites"] += 1
            self.stats["pushed_bytes"] += sum(ps.nbytes() for ps in mine_out)
            self.stats["received_bytes"] += sum(ps.nbytes() for ps in mine_in)
            if getattr(self.transport, "split_sites", False) and self.backend.recording is not None:
                # Split site: the PUSH (copy my rows into the consumers' buffers + raise their flags) is issued right behind the
                # launch that produced the rows - it then overlaps whatever this rank computes next - and only the WAIT for my
                # own incoming rows stays in front of the consumer.  Records carry order keys; _reorder() sorts them once.
                self._sync_keys()
                if mine_out:
                    self.transport.push(self.site, mine_out)
                    self._sync_keys(max(ps.buf.ready_key for ps in mine_out) + 0.5)
                if mine_in:
                    self.transport.wait(self.site, mine_in)
                    k = self._sync_keys()
                    for ps in mine_in:
                        ps.buf.ready_key = max(ps.buf.ready_key, k)
            else:
                self.transport.exchange(self.site, mine_out, mine_in)
            self.site += 1
        self._sync_keys()
        launch()
        k_launch = self._sync_keys()
        for p in range(self.world):
            for view, rows in writes[p]:
                if not rows:
                    continue
                b, i0, ni = self._buf(view)
                b.external = False
                if p == self.rank:
                    b.ready_key = max(b.ready_key, k_launch)
                for img in range(i0, i0 + ni):
                    b.have[p][img] = rs_union(b.have[p][img], rows)

    @staticmethod
    def _merge(pushes: List[Push]) -> List[Push]:
        """Merge pushes of the same rows of consecutive images (one piece with several chunks)."""
        pushes.sort(key=lambda s: (s.src, s.dst, s.buf.idx, s.lo, s.hi, s.img0))
        out: List[Push] = []
        for s in pushes:
            l = out[-1] if out else None
            if l is not None and (l.src, l.dst, l.buf, l.lo, l.hi) == (s.src, s.dst, s.buf, s.lo, s.hi) and l.img0 + l.nimg == s.img0:
                l.nimg += 1
            else:
                out.append(s)
        return out

    def _iv(self, iv: Interval) -> RowSet:
        return [iv] if iv[1] > iv[0] else []

    def _all(self, x) -> RowSet:
        return [(0, self._buf(x)[0].H)]

    def _per_rank(self, fn):
        reads, writes = [], []
        for p in range(self.world):
            r, w = fn(p)
            reads.append(r)
            writes.append(w)
        return reads, writes

    # ---- GEMM-shaped layers -----------------------------------------------------------------------
    def gemm_conv(self, srcs: Sequence[Map], w: PackedGemm, out: Map, *, stride=1, dil=1, act=True, residual=None, out2=None,
                  prelu2=None, win: Optional[WinGeom] = None, precision=None, qkv_heads: int = 0, out_f32: bool = False):
        ob = self._buf(out)[0]
        sb = self._buf(srcs[0])[0]
        gsrcs = [self._buf(s)[0].grid_map(s) for s in srcs]
        gout = ob.grid_map(out)
        gres = None if residual is None else self._buf(residual)[0].grid_map(residual)
        gout2 = None if out2 is None else self._buf(out2)[0].grid_map(out2)
        k = w.ksize
        pad = dil * (k - 1) // 2
        Hin = sb.H

        def gemm_rows(p: int) -> Interval:
            if win is not None:
                assert sb.kind == "win"
                return self.own(sb, p)
            a, b = self.own(ob, p)
            if w.shuffle:
                return (a // 2, (b + 1) // 2)
            return (a, b)

        def deps(p: int):
            a, b = gemm_rows(p)
            if b <= a:
                return [], []
            src_rows = [(max(0, a * stride - pad), min(Hin, (b - 1) * stride - pad + dil * (k - 1) + 1))]
            reads = [(s, src_rows) for s in srcs]
            if win is not None:
                out_rows = win_rows_to_tokens(win, a, b)
                if residual is not None:
                    reads.append((residual, [(a, b)]))          # residual is window-major like the GEMM rows
            else:
                out_rows = [(2 * a, 2 * b)] if w.shuffle else [(a, b)]
                if residual is not None:
                    reads.append((residual, out_rows))
            writes = [(out, out_rows)] + ([(out2, out_rows)] if out2 is not None else [])
            return reads, writes

        reads, writes = self._per_rank(deps)
        mine = gemm_rows(self.rank)

        def launch():
            if mine[1] > mine[0]:
                self.backend.gemm_conv(gsrcs, w, gout, stride=stride, dil=dil, act=act, residual=gres, out2=gout2, prelu2=prelu2,
                                       win=win, precision=precision, rows=mine, qkv_heads=qkv_heads)

        self._run(reads, writes, launch)

    def _same_rows(self, outs, ins, call, all_rows_in=()):
        """Per-pixel op: outputs and ``ins`` on the same rows; ``all_rows_in`` (warp sources) are needed whole."""
        ob = self._buf(outs[0])[0]

        def deps(p):
            iv = self._iv(self.own(ob, p))
            if not iv:
                return [], []
            return [(x, iv) for x in ins] + [(x, self._all(x)) for x in all_rows_in], [(x, iv) for x in outs]

        reads, writes = self._per_rank(deps)
        mine = self.own(ob, self.rank)
        self._run(reads, writes, (lambda: call(mine)) if mine[1] > mine[0] else (lambda: None))

    def conv3x3_first(self, img, w: PackedGemm, out: Map):
        ob = self._buf(out)[0]

        def deps(p):
            a, b = self.own(ob, p)
            if b <= a:
                return [], []
            return [(img, [(max(0, a - 1), min(ob.H, b + 1))])], [(out, [(a, b)])]

        reads, writes = self._per_rank(deps)
        mine = self.own(ob, self.rank)
        self._run(reads, writes, lambda: self.backend.conv3x3_first(img, w, out, rows=mine) if mine[1] > mine[0] else None)

    def pack5_planar(self, imgs, out: Map):
        self._same_rows([out], list(imgs), lambda rows: self.backend.pack5_planar(imgs, out, rows=rows))

    # ---- transformer pieces -----------------------------------------------------------------------
    def layernorm(self, x: Map, out: Map, gamma, beta):
        gx, go = self._buf(x)[0].grid_map(x), self._buf(out)[0].grid_map(out)
        self._same_rows([out], [x], lambda rows: self.backend.layernorm(gx, go, gamma, beta, rows=rows))

    def window_gather_ln(self, tok: Map, win: Map, g: WinGeom, gamma, beta):
        wb = self._buf(win)[0]

        def deps(p):
            k0, k1 = self.own(wb, p)
            if k1 <= k0:
                return [], []
            return [(tok, win_rows_to_tokens(g, k0, k1))], [(win, [(k0, k1)])]

        reads, writes = self._per_rank(deps)
        mine = self.own(wb, self.rank)
        self._run(reads, writes, lambda: self.backend.window_gather_ln(tok, win, g, gamma, beta, rows=mine) if mine[1] > mine[0] else None)

    def window_attention(self, qkv: Map, out: Map, g: WinGeom, heads, cross, rc=None, mix=None, motion=None, motion_off=0,
                         scratch=None, rc_closed_form=False, head_major=False):
        qb = self._buf(qkv)[0]

        def deps(p):
            k0, k1 = self.own(qb, p)
            if k1 <= k0:
                return [], []
            writes = [(out, [(k0, k1)])]
            if motion is not None:
                writes.append((motion, win_rows_to_tokens(g, k0, k1)))
            return [(qkv, [(k0, k1)])], writes

        reads, writes = self._per_rank(deps)
        mine = self.own(qb, self.rank)
        if head_major:
            # Q[h][r][d] / K / V^T planes: a row of windows is not one contiguous chunk of this buffer, so it must never be
            # exchanged - and it is not: the rank that owns a row of windows produced its q | k | v itself
            for p in range(self.world):
                for view, need in reads[p]:
                    b, i0, ni = self._buf(view)
                    assert all(not rs_sub(need, b.have[p][i]) for i in range(i0, i0 + ni)), "head-major qkv rows would have to be exchanged"

        def launch():
            if mine[1] > mine[0]:
                self.backend.window_attention(qkv, out, g, heads, cross, rc, mix, motion, motion_off, scratch,
                                              rc_closed_form=rc_closed_form, rows=mine, head_major=head_major)

        self._run(reads, writes, launch)

    def dwconv_gelu(self, x: Map, out: Map, w9c, bias):
        ob = self._buf(out)[0]

        def deps(p):
            a, b = self.own(ob, p)
            if b <= a:
                return [], []
            return [(x, [(max(0, a - 1), min(ob.H, b + 1))])], [(out, [(a, b)])]

        reads, writes = self._per_rank(deps)
        mine = self.own(ob, self.rank)
        self._run(reads, writes, lambda: self.backend.dwconv_gelu(x, out, w9c, bias, rows=mine) if mine[1] > mine[0] else None)

    def mlp_tail_ok(self, *a) -> bool:
        f = getattr(self.backend, "mlp_tail_ok", None)
        return bool(f and f(*a))

    def mlp_tail(self, hidden: Map, dw_w, dw_b, fc2, residual: Map, out: Map):
        """Fused DWConv + GELU + fc2 + residual: the depth-wise taps reach one hidden row beyond the slab on either side."""
        ob = self._buf(out)[0]

        def deps(p):
            a, b = self.own(ob, p)
            if b <= a:
                return [], []
            return [(hidden, [(max(0, a - 1), min(ob.H, b + 1))]), (residual, [(a, b)])], [(out, [(a, b)])]

        reads, writes = self._per_rank(deps)
        mine = self.own(ob, self.rank)
        self._run(reads, writes, lambda: self.backend.mlp_tail(hidden, dw_w, dw_b, fc2, residual, out, rows=mine) if mine[1] > mine[0] else None)

    # ---- warps, resampling, layout ------------------------------------------------------------------
    def flow_warp_nchw(self, img, flow, out):
        self._same_rows([out], [flow], lambda rows: self.backend.flow_warp_nchw(img, flow, out, rows=rows), all_rows_in=[img])

    def flow_warp_nhwc(self, src: Map, head: Map, flow_off: int, out: Map):
        if not getattr(self.transport, "remote_reads", False) or self._replicated:
            # transports without peer-mapped memory: gather the whole source in front of the warp
            self._same_rows([out], [head], lambda rows: self.backend.flow_warp_nhwc(src, head, flow_off, out, rows=rows), all_rows_in=[src])
            return
        # NVLink: the source stays where it was produced; the kernel reads each sample from the GPU that owns the row.
        # One flag-only site makes sure every rank has finished producing its rows.
        owners = self._owners(src)
        self._sync_site()
        self._same_rows([out], [head], lambda rows: self.backend.flow_warp_nhwc(src, head, flow_off, out, rows=rows, owners=owners))

    def _owners(self, x):
        """Cover of the rows of ``x`` by (lo, hi, byte distance to the rank that holds them), preferring the local copy; None
        when every row is local (external inputs, replicated buffers)."""
        b, i0, ni = self._buf(x)
        if b.external:
            return None
        segs = []
        todo = [(0, b.H)]
        for q in [self.rank] + [q for q in range(self.world) if q != self.rank]:
            have = b.have[q][i0]
            assert all(b.have[q][i] == have for i in range(i0, i0 + ni)), "images of one view must share their row layout"
            got = rs_and(todo, have)
            segs += [(lo, hi, q) for lo, hi in got]
            todo = rs_sub(todo, got)
        assert not todo, f"rows {todo} of the warp source exist on no rank"
        segs.sort()
        if all(q == self.rank for _, _, q in segs):
            return None
        return [(lo, hi, self.transport.byte_delta(q)) for lo, hi, q in segs]

    def _sync_site(self):
        """Flag-only exchange site: every rank has finished producing what its peers are about to read in place."""
        self.transport.barrier(self.site)
        self.site += 1
        self.stats["sites"] += 1

    def warp_blend(self, im0, im1, head: Map, w0, w1, it, flow0=None, flow1=None, occ1=None, occ2=None):
        outs = [t for t in (w0, w1, it, flow0, flow1, occ1, occ2) if t is not None]
        if getattr(self.transport, "remote_reads", False) and not self._replicated:
            # the (per-slab) warped pyramids are read in place from their owners; whether a source is local is the same on all ranks
            o0, o1 = self._owners(im0), self._owners(im1)
            if o0 is not None or o1 is not None:
                b0, b1 = self._buf(im0)[0], self._buf(im1)[0]
                assert o0 is not None and o1 is not None and b0.have == b1.have, "the two warp sources must share their row layout"
                self._sync_site()
                self._same_rows(outs, [head], lambda rows: self.backend.warp_blend(im0, im1, head, w0, w1, it, flow0, flow1, occ1, occ2,
                                                                                   rows=rows, owners=o0))
                return
        self._same_rows(outs, [head], lambda rows: self.backend.warp_blend(im0, im1, head, w0, w1, it, flow0, flow1, occ1, occ2, rows=rows),
                        all_rows_in=[im0, im1])

    def resize(self, x, out, scale: float = 1.0):
        ob = self._buf(out)[0]
        Hin, Hout = x.shape[2], out.shape[2]
        sh = np.float32(Hin - 1) / np.float32(Hout - 1) if Hout > 1 else np.float32(0)

        def src_rows(a, b):     # rows touched by the bilinear taps of output rows [a, b) (csrc/elementwise.cu resize_ac_kernel)
            lo = int(np.float32(sh * np.float32(a)))
            hi = int(np.float32(sh * np.float32(b - 1))) + 2
            return [(max(0, lo), min(Hin, hi))]

        def deps(p):
            a, b = self.own(ob, p)
            if b <= a:
                return [], []
            return [(x, src_rows(a, b))], [(out, [(a, b)])]

        reads, writes = self._per_rank(deps)
        mine = self.own(ob, self.rank)
        self._run(reads, writes, lambda: self.backend.resize(x, out, scale, rows=mine) if mine[1] > mine[0] else None)

    def nchw_to_nhwc(self, x, out: Map, zero_fill_to: int = 0):
        self._same_rows([out], [x], lambda rows: self.backend.nchw_to_nhwc(x, out, zero_fill_to, rows=rows))

    def residual_finish(self, res: Map, it, it_sum, it_clamped):
        outs = [t for t in (it_sum, it_clamped) if t is not None]
        self._same_rows(outs, [res, it], lambda rows: self.backend.residual_finish(res, it, it_sum, it_clamped, rows=rows))

    # ---- results ------------------------------------------------------------------------------------
    def gather_outputs(self, outputs: Dict[str, object]) -> None:
        """Rank 0 receives every rank's rows of the public output tensors (``gather``: all of them, only I_t, or none)."""
        if self.gather == "none":
            return
        tensors: List[torch.Tensor] = []
        for key, v in outputs.items():
            if self.gather == "I_t" and key != "I_t":
                continue
            for t in (v if isinstance(v, list) else [v]):
                if t is not None and all(t is not u for u in tensors):
                    tensors.append(t)
        reads = [[(t, self._all(t)) for t in tensors] if p == 0 else [] for p in range(self.world)]
        self._run(reads, [[] for _ in range(self.world)], lambda: None)
        self._reorder()
