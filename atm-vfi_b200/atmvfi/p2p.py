"""NVLink peer-to-peer plumbing of the row-slab mode: an IPC-shared arena per rank and the exchange transport
that ``slab.SlabOps`` schedules against (C side: csrc/p2p.cu, ABI: include/atmvfi.h "Spatial row-slab mode").

One process per GPU (torchrun).  ``torch.distributed`` is used ONLY to pass the 64-byte CUDA IPC handles around
and for the host-side barrier after set-up; rows move through peer-mapped stores issued by our own kernel.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional

import torch
import torch.distributed as dist

from . import _lib
from .engine import PackedModel, Plan
from .ops import CudaOps
from .slab import Push, SlabOps

CTRL_BYTES = 1 << 20            # control block at the start of every arena
OFF_EPOCH, OFF_ERROR, OFF_EPOCH_IN, OFF_READY, OFF_READY_IN, OFF_COUNTERS, OFF_FLAGS = 0, 16, 32, 64, 128, 4096, 65536
MAX_SITES = 8192
INPUT_SITE = MAX_SITES - 1      # exchange site of the input-frame all-gather (its flags count the input epoch OFF_EPOCH_IN)
ALIGN = 1024


class _DevMem:
    """Raw device memory exposed to torch through __cuda_array_interface__ (no copy, no ownership)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


class P2PArena:
    """One cudaMalloc'ed block per rank with the same layout everywhere; peers' blocks are mapped through CUDA IPC."""

    def __init__(self, device: torch.device, nbytes: int, group=None):
        self.lib = _lib.load()
        self.device, self.group = device, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > _lib.P2P_MAX_PEERS:
            raise _lib.AtmvfiError(f"row slabs support up to {_lib.P2P_MAX_PEERS} ranks, got {self.world}")
        self.nbytes = (nbytes + ALIGN - 1) // ALIGN * ALIGN
        with torch.cuda.device(device):
            p = C.c_void_p()
            _lib.check(self.lib.atmvfi_arena_alloc(self.nbytes, C.byref(p)), "arena_alloc")
            self.base = p.value
            self.bytes_view = torch.as_tensor(_DevMem(self.base, self.nbytes), device=device)
            assert self.bytes_view.data_ptr() == self.base, "torch copied the arena instead of wrapping it"
            self.bytes_view.zero_()
            torch.cuda.synchronize(device)
            handle = C.create_string_buffer(_lib.IPC_HANDLE_BYTES)
            _lib.check(self.lib.atmvfi_ipc_export(self.base, handle), "ipc_export")
            handles: List[Optional[bytes]] = [None] * self.world
            dist.all_gather_object(handles, bytes(handle.raw), group=group)
            self.peer_base: List[int] = []
            for r, h in enumerate(handles):
                if r == self.rank:
                    self.peer_base.append(self.base)
                    continue
                q = C.c_void_p()
                _lib.check(self.lib.atmvfi_ipc_open(C.create_string_buffer(h, _lib.IPC_HANDLE_BYTES), C.byref(q)), f"ipc_open(rank {r})")
                self.peer_base.append(q.value)
            dist.barrier(group=group)                  # every arena is zeroed and mapped before anyone pushes
        self.top = CTRL_BYTES
        self._closed = False

    def alloc(self, shape, zero: bool) -> torch.Tensor:
        n = 4
        for s in shape:
            n *= int(s)
        off = self.top
        self.top = (off + n + ALIGN - 1) // ALIGN * ALIGN
        if self.top > self.nbytes:
            raise _lib.AtmvfiError(f"row-slab arena of {self.nbytes >> 20} MiB exhausted (need > {self.top >> 20} MiB); pass a larger arena_bytes")
        # the arena starts zeroed and is never recycled, so `zero` needs no extra work
        return self.bytes_view[off : off + n].view(torch.float32).view(*shape)

    def peer(self, rank: int, local_ptr: int) -> int:
        off = local_ptr - self.base
        assert 0 <= off < self.nbytes, "pointer is not inside the arena"
        return self.peer_base[rank] + off

    def ctrl(self, rank: int, offset: int) -> int:
        return self.peer_base[rank] + offset

    def error_flag(self) -> int:
        return int(self.bytes_view[OFF_ERROR : OFF_ERROR + 4].view(torch.int32).item())

    def reset_control(self) -> None:
        """Zero the control block (epoch, ready flags, site counters and flags, error word).  Collective: every rank calls it."""
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)                 # nobody is still spinning on / pushing into a control block
        self.bytes_view[:CTRL_BYTES].zero_()
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)

    def close(self) -> None:
        if self._closed:
            return
        self._closed = True
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)
        for r, p in enumerate(self.peer_base):
            if r != self.rank:
                self.lib.atmvfi_ipc_close(p)
        dist.barrier(group=self.group)
        self.bytes_view = None
        self.lib.atmvfi_arena_free(self.base)


class P2PTransport:
    """``slab.SlabOps`` transport: one ``atmvfi_p2p_exchange`` launch per exchange site."""

    def __init__(self, arena: P2PArena):
        self.arena, self.rank, self.world = arena, arena.rank, arena.world
        self.slab: Optional[SlabOps] = None

    def attach(self, slab: SlabOps) -> None:
        self.slab = slab

    def _ptr_array(self, ptrs: List[int]):
        return (C.c_void_p * max(1, len(ptrs)))(*ptrs)

    def step_begin(self) -> None:
        a = self.arena
        others = [r for r in range(self.world) if r != self.rank]
        sig = self._ptr_array([a.ctrl(r, OFF_READY + 4 * self.rank) for r in others])
        wait = self._ptr_array([a.ctrl(self.rank, OFF_READY + 4 * r) for r in others])
        self.slab.backend._emit("atmvfi_p2p_step_begin", (a.ctrl(self.rank, OFF_EPOCH), sig, len(others), wait, len(others), a.ctrl(self.rank, OFF_ERROR)),
                                keep=(sig, wait))

    remote_reads = True            # kernels may dereference peer-mapped addresses (flow_warp_nhwc_p2p)
    split_sites = True             # SlabOps issues push and wait of an exchange site separately (push early, wait late)

    def byte_delta(self, rank: int) -> int:
        """Distance from a local arena address to the same buffer in ``rank``'s arena (peer-mapped address space)."""
        return self.arena.peer_base[rank] - self.arena.base

    def barrier(self, site: int) -> None:
        """Flag-only site: every rank signals every other rank and waits for all of them."""
        a, me = self.arena, self.rank
        others = [r for r in range(self.world) if r != me]
        if not others:
            return
        sig = self._ptr_array([a.ctrl(d, OFF_FLAGS + 4 * (site * _lib.P2P_MAX_PEERS + me)) for d in others])
        wait = self._ptr_array([a.ctrl(me, OFF_FLAGS + 4 * (site * _lib.P2P_MAX_PEERS + s)) for s in others])
        arr = (_lib.P2PPiece * 1)()
        self.slab.backend._emit("atmvfi_p2p_exchange", (arr, 0, sig, len(others), wait, len(others), a.ctrl(me, OFF_EPOCH),
                                                        a.ctrl(me, OFF_COUNTERS + 4 * site), a.ctrl(me, OFF_ERROR)), keep=(arr, sig, wait))

    def push(self, site: int, outgoing: List[Push]) -> None:
        """First half of a split site: copy my rows into the consumers' buffers and raise their flags; waits for nobody."""
        self.exchange(site, outgoing, [])

    def wait(self, site: int, incoming: List[Push]) -> None:
        """Second half: spin (on the device) until every producer of my incoming rows has raised this site's flag."""
        self.exchange(site, [], incoming)

    def gather_inputs(self, frames: List[torch.Tensor], bounds: List[int]) -> None:
        """All-gather of the input frames over NVLink: every rank holds rows [bounds[r], bounds[r+1]) of each planar frame
        [1, C, H, W] (it converted them from ITS share of the uploaded uint8 rows) and pushes them to every peer.  A barrier on a
        second epoch word comes first: a peer may still be reading the previous pair's frames.  Launched eagerly, in front of the
        captured plan; flags of the input site carry the input epoch."""
        a, me = self.arena, self.rank
        others = [r for r in range(self.world) if r != me]
        if not others:
            return
        emit = self.slab.backend._emit
        sig = self._ptr_array([a.ctrl(r, OFF_READY_IN + 4 * me) for r in others])
        wait = self._ptr_array([a.ctrl(me, OFF_READY_IN + 4 * r) for r in others])
        emit("atmvfi_p2p_step_begin", (a.ctrl(me, OFF_EPOCH_IN), sig, len(others), wait, len(others), a.ctrl(me, OFF_ERROR)), keep=(sig, wait))
        lo, hi = bounds[me], bounds[me + 1]
        pieces = []
        for t in frames:
            _, c, h, w = t.shape
            start = t.data_ptr() + lo * w * 4
            for d in others:
                pieces.append(_lib.P2PPiece(start, a.peer(d, start), (hi - lo) * w * 4, h * w * 4, c, 0))
        site = INPUT_SITE
        sig = self._ptr_array([a.ctrl(d, OFF_FLAGS + 4 * (site * _lib.P2P_MAX_PEERS + me)) for d in others])
        wait = self._ptr_array([a.ctrl(me, OFF_FLAGS + 4 * (site * _lib.P2P_MAX_PEERS + s)) for s in others])
        n = _lib.P2P_MAX_PIECES
        groups = [pieces[i : i + n] for i in range(0, len(pieces), n)]
        for gi, grp in enumerate(groups):
            arr = (_lib.P2PPiece * len(grp))(*grp)
            last = gi == len(groups) - 1
            emit("atmvfi_p2p_exchange", (arr, len(grp), sig, len(others) if last else 0, wait, len(others) if last else 0, a.ctrl(me, OFF_EPOCH_IN),
                                         a.ctrl(me, OFF_COUNTERS + 4 * site), a.ctrl(me, OFF_ERROR)), keep=(arr, sig, wait))

    def exchange(self, site: int, outgoing: List[Push], incoming: List[Push]) -> None:
        if not outgoing and not incoming:
            return
        if site >= INPUT_SITE:
            raise _lib.AtmvfiError(f"more than {INPUT_SITE} exchange sites in one plan")
        a, me = self.arena, self.rank
        pieces = []
        for ps in outgoing:
            b = ps.buf
            start = b.t.data_ptr() + ((ps.img0 * b.planes) * b.H + ps.lo) * b.row_bytes
            pc = _lib.P2PPiece(start, a.peer(ps.dst, start), (ps.hi - ps.lo) * b.row_bytes, b.H * b.row_bytes, ps.nimg * b.planes, 0)
            pieces.append(pc)
        dsts = sorted({ps.dst for ps in outgoing})
        srcs = sorted({ps.src for ps in incoming})
        sig = self._ptr_array([a.ctrl(d, OFF_FLAGS + 4 * (site * _lib.P2P_MAX_PEERS + me)) for d in dsts])
        wait = self._ptr_array([a.ctrl(me, OFF_FLAGS + 4 * (site * _lib.P2P_MAX_PEERS + s)) for s in srcs])
        epoch, counter, err = a.ctrl(me, OFF_EPOCH), a.ctrl(me, OFF_COUNTERS + 4 * site), a.ctrl(me, OFF_ERROR)
        n = _lib.P2P_MAX_PIECES
        groups = [pieces[i : i + n] for i in range(0, len(pieces), n)] or [[]]
        for gi, grp in enumerate(groups):
            arr = (_lib.P2PPiece * max(1, len(grp)))(*grp)
            last = gi == len(groups) - 1
            self.slab.backend._emit("atmvfi_p2p_exchange",
                                    (arr, len(grp), sig, len(dsts) if last else 0, wait, len(srcs) if last else 0, epoch, counter, err),
                                    keep=(arr, sig, wait, [ps.buf.t for ps in outgoing]))


def default_arena_bytes(arch_name: str, B: int, H: int, W: int) -> int:
    """Plan buffers measured at 19 GB (Base, 1088x1920) / 78 GB (Base, 2176x4096): ~9.2 KB per pixel; Lite is ~half."""
    per_px = 10.5e3 if arch_name == "base" else 6.5e3
    return int(per_px * B * H * W) + (768 << 20)


class SlabSession:
    """One frame pair (or batch) split into row slabs over the ranks of ``group``.  Every rank calls ``run`` with the FULL
    frames; rank 0 returns the gathered outputs (``gather``: "I_t" only, "all" ten entries, "none")."""

    def __init__(self, net, B: int, H: int, W: int, global_motion: Optional[bool] = None, group=None, gather: str = "I_t",
                 arena_bytes: Optional[int] = None, timeout_ms: Optional[int] = None):
        from .runtime import PRECISIONS
        self._failed = False
        import os
        self.sliced_upload = os.environ.get("ATMVFI_SLAB_SLICED_UPLOAD", "1") != "0"
        if timeout_ms is not None:          # peer-wait time-out of the exchange kernels (default 4 s / ATMVFI_P2P_TIMEOUT_MS)
            _lib.check(_lib.load().atmvfi_p2p_set_timeout_ms(int(timeout_ms)), "p2p_set_timeout_ms")
        dev = next(net.parameters()).device
        if dev.type != "cuda":
            raise _lib.AtmvfiError("row slabs need the model on a CUDA device; there is no CPU fallback")
        if not dist.is_initialized():
            raise _lib.AtmvfiError("row slabs need torch.distributed (one process per GPU, launched with torchrun)")
        glob = bool(net.global_motion if global_motion is None else global_motion)
        if net.precision == "f16":
            raise _lib.AtmvfiError("row slabs exchange fp32 feature-map rows: use precision 'tf32', 'fp32x3' or 'fp32' (not 'f16')")
        self.device = dev
        with torch.cuda.device(dev):
            self.arena = P2PArena(dev, arena_bytes or default_arena_bytes(net.ARCH.name, B, H, W), group)
            self.ops = CudaOps(dev, PRECISIONS[net.precision])
            self.ops.allocator = self.arena.alloc
            self.transport = P2PTransport(self.arena)
            self.slab = SlabOps(self.ops, self.arena.rank, self.arena.world, self.transport, gather)
            sd = {k: v.detach() for k, v in net.state_dict().items()}
            self.model = PackedModel(net.ARCH, sd, net.local_motion_args["window_size"], net.global_motion_args["window_size"], with_global=True)
            self.plan = Plan(self.slab, self.model, B, H, W, glob)
            torch.cuda.synchronize(dev)
            dist.barrier(group=group)
        self.rank, self.world = self.arena.rank, self.arena.world

    def _usable(self) -> None:
        if self._failed:
            raise _lib.AtmvfiError("this row-slab session saw a peer time-out; call resync() on every rank (or rebuild the session)")

    def run(self, im0: torch.Tensor, im1: torch.Tensor, use_graph: bool = True, check: bool = True) -> Dict[str, object]:
        """``check=True`` (default) waits for the step and raises if a peer wait timed out - a poisoned step returns garbage rows.
        Pipelined callers pass ``check=False`` and call ``check()`` themselves before they trust the frames."""
        self._usable()
        with torch.cuda.device(self.device):
            out = self.plan.run(im0, im1, use_graph=use_graph)
        if check:
            self.check()
        return out

    def run_inplace(self, use_graph: bool = True, check: bool = True) -> Dict[str, object]:
        self._usable()
        with torch.cuda.device(self.device):
            out = self.plan.run_inplace(use_graph=use_graph)
        if check:
            self.check()
        return out

    def _stage(self, H: int, W: int) -> dict:
        st = getattr(self, "_staging", None)
        if st is None or tuple(st["h0"].shape[:2]) != (H, W):
            mk_h = lambda: torch.empty((H, W, 3), dtype=torch.uint8).pin_memory()
            mk_d = lambda: torch.empty((H, W, 3), dtype=torch.uint8, device=self.device)
            st = self._staging = dict(h0=mk_h(), h1=mk_h(), hout=mk_h(), d0=mk_d(), d1=mk_d(), dout=mk_d())
        return st

    def pinned_frame_buffers(self, H: int, W: int):
        """Two HxWx3 uint8 numpy arrays in pinned memory; frames decoded straight into them are uploaded without a host copy."""
        st = self._stage(H, W)
        return st["h0"].numpy(), st["h1"].numpy()

    def interpolate_u8(self, img0, img1, isBGR: bool = True, divisor: int = 64, copy: bool = False):
        """``demo_2x.inference_2frame`` arithmetic (demo_2x.py:54-87) on the slab plan: every rank passes the same two HxWx3
        uint8 frames; rank 0 returns the uint8 middle frame, the other ranks return None.  With ``copy=False`` (default) the
        result is a view of the session's pinned download buffer, valid until the next call."""
        H, W = img0.shape[:2]
        eh, ew = (-H) % divisor, (-W) % divisor
        Hp, Wp, top, left = H + eh, W + ew, eh // 2, ew // 2
        assert tuple(self.plan.im0.shape) == (1, 3, Hp, Wp), "session was built for another frame size"
        self._usable()
        with torch.cuda.device(self.device):
            st = self._stage(H, W)
            if self.sliced_upload and self.world > 1:
                # every rank uploads only the uint8 rows behind ITS slab of the padded frame (1/N of the PCIe traffic), converts
                # them, and the fp32 planar rows are all-gathered over NVLink (P2PTransport.gather_inputs)
                b = self.slab.bounds
                y0, y1 = b[self.rank], b[self.rank + 1]
                s0, s1 = min(max(y0 - top, 0), H - 1), min(max(y1 - 1 - top, 0), H - 1) + 1
                for src, h, d, dst in ((img0, st["h0"], st["d0"], self.plan.im0), (img1, st["h1"], st["d1"], self.plan.im1)):
                    if src.__array_interface__["data"][0] != h.data_ptr():
                        h.numpy()[s0:s1] = src[s0:s1]
                    d[s0:s1].copy_(h[s0:s1], non_blocking=True)
                    self.ops._emit("atmvfi_u8_to_planar_rows", (d.data_ptr(), dst.data_ptr(), H, W, Hp, Wp, top, left, int(isBGR), y0, y1), keep=(d, dst))
                self.transport.gather_inputs([self.plan.im0, self.plan.im1], b)
                self.h2d_bytes_this_rank = 2 * (s1 - s0) * W * 3
            else:
                for src, h, d in ((img0, st["h0"], st["d0"]), (img1, st["h1"], st["d1"])):
                    if src.__array_interface__["data"][0] != h.data_ptr():      # not already in the pinned buffers
                        h.numpy()[...] = src
                    d.copy_(h, non_blocking=True)
                self.ops.u8_to_planar(st["d0"], self.plan.im0, H, W, Hp, Wp, top, left, isBGR)
                self.ops.u8_to_planar(st["d1"], self.plan.im1, H, W, Hp, Wp, top, left, isBGR)
                self.h2d_bytes_this_rank = 2 * H * W * 3
            out = self.plan.run_inplace(use_graph=True)
            if self.rank == 0:
                self.ops.planar_to_u8(out["I_t"], st["dout"], H, W, Hp, Wp, top, left, isBGR)
                st["hout"].copy_(st["dout"], non_blocking=True)
            torch.cuda.current_stream().synchronize()
            self.check()                     # the stream is idle: reading the 4-byte error word costs one tiny D2H copy
            if self.rank != 0:
                return None
            return st["hout"].numpy().copy() if copy else st["hout"].numpy()

    def check(self) -> None:
        """Raise if a peer wait timed out on the device (a rank died, stalled longer than the time-out, or the ranks ran
        different step counts).  The step that timed out and every later one pushed nothing, so their frames are garbage; the
        session refuses further work until ``resync()``."""
        torch.cuda.synchronize(self.device)
        if self.arena.error_flag():
            self._failed = True
            raise _lib.AtmvfiError("row-slab exchange timed out waiting for a peer (time-out: atmvfi_p2p_set_timeout_ms / "
                                   "ATMVFI_P2P_TIMEOUT_MS); the frames of this step are invalid")

    def resync(self) -> None:
        """Collective recovery after a time-out: all ranks reset their control blocks (epochs, flags, error word) together."""
        self.arena.reset_control()
        self._failed = False

    def close(self) -> None:
        self.plan = None
        self.arena.close()
