"""Shared implementation of ``network_base.Network`` and ``network_lite.Network``.

Same public surface as the reference classes (network/network_base.py:88-340): constructor arguments,
the mutable ``global_motion`` / ``ensemble_global_motion`` attributes, the window-size setters, the
freeze / finetune toggles, the 236-entry state-dict and the ``forward(im0, im1) -> dict`` contract.
The forward itself is the sm_100a engine (atmvfi/engine.py); there is no PyTorch or CPU fallback.
"""
from __future__ import annotations

import os
import sys

import torch
import torch.nn as nn

_PKG_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _PKG_ROOT not in sys.path:
    sys.path.insert(0, _PKG_ROOT)

from atmvfi import arch as _arch                      # noqa: E402
from atmvfi.engine import clone_outputs               # noqa: E402
from atmvfi.modules import ParamTree, relative_coord_buffer   # noqa: E402
from atmvfi.runtime import Runtime                    # noqa: E402

_LOCAL_PARTS = ("feat_extracts", "cross_scale_feature_fusion", "local_motion_atmformer", "local_motion_mlp",
                "feat_enhance_transformer", "upsample_pyramid", "proj", "down1", "down2", "down3", "up1", "up2", "up3",
                "refine_head")
_GLOBAL_PARTS = ("last_feat_extract", "global_feature_fusion", "global_motion_atmformer", "global_motion_mlp")
_REFINE_PARTS = ("proj", "down1", "down2", "down3", "up1", "up2", "up3", "refine_head")


class NetworkBase(ParamTree):
    ARCH: _arch.Arch = None

    def __init__(self, global_motion=True, ensemble_global_motion=False):
        super().__init__()
        a = self.ARCH
        self.pyramid_level = 4
        self.hidden_dims = list(a.enc)
        self.global_motion = global_motion
        self.ensemble_global_motion = ensemble_global_motion
        self.local_motion_args = {"window_size": 8, "num_heads": _arch.NUM_HEADS, "patch_size": 1, "dim": a.C,
                                  "enhance_window": _arch.ENHANCE_WINDOW}
        self.global_motion_args = {"window_size": 12, "num_heads": _arch.NUM_HEADS, "patch_size": 1, "dim": a.GC}
        if a.name == "lite":
            self.local_motion_args["mlp_ratio"] = a.mlp_ratio
            self.global_motion_args["mlp_ratio"] = a.mlp_ratio
        self.fused_dim = 2 * a.C
        self.motion_out_dim = _arch.MOTION_OUT
        self.fused_dim1, self.fused_dim2, self.fused_dim3 = a.C, a.C // 2, a.C // 4
        self.fused_dims = [self.fused_dim1, self.fused_dim2, self.fused_dim3, 2 * self.fused_dim1]
        self.populate(_arch.param_schema(a))
        # engine knobs (not part of the reference surface)
        self.precision = os.environ.get("ATMVFI_PRECISION", "tf32")   # "tf32": tcgen05 kind::tf32; "fp32": CUDA-core FFMA
        self.use_cuda_graph = os.environ.get("ATMVFI_CUDA_GRAPH", "1") != "0"
        self.zero_copy_outputs = False
        # interpolate_stream: encode every interior frame once (it is frame 1 of one pair and frame 0 of the next)
        self.stream_encoder_reuse = os.environ.get("ATMVFI_STREAM_REUSE", "1") != "0"
        self._runtime = Runtime(a)

    # ---- reference API: window sizes (network_base.py:262-270) -------------------------------------
    def _set_ws(self, blocks: ParamTree, ws: int) -> None:
        for k in ("0", "1"):
            attn = blocks._modules[k]._modules["attn"]
            dev = attn.relative_coord.device
            attn.relative_coord = relative_coord_buffer(ws).to(dev)

    def __set_local_window_size__(self, window_size):
        self.local_motion_args["window_size"] = window_size
        self._set_ws(self.local_motion_atmformer, window_size)

    def __set_global_window_size__(self, window_size):
        self.global_motion_args["window_size"] = window_size
        self._set_ws(self.global_motion_atmformer, window_size)

    # ---- reference API: training toggles (network_base.py:272-334); names kept so callers resolve ---
    def _grad(self, parts, flag):
        for p in parts:
            getattr(self, p).requires_grad_(flag)

    def __freeze_global_motion__(self):
        self._grad(_GLOBAL_PARTS, False)

    def __finetune_global_motion__(self):
        self._grad(_GLOBAL_PARTS, True)

    def __freeze_local_motion__(self):
        self._grad(_LOCAL_PARTS, False)

    def __finetune_local_motion__(self):
        self._grad(_LOCAL_PARTS, True)

    # ---- forward ------------------------------------------------------------------------------------
    def forward(self, im0, im1):
        if self.ensemble_global_motion:
            return self.forward_global_ensemble(im0, im1)
        return self.forward_normal(im0, im1)

    def _check_inputs(self, im0, im1):
        if im0.dim() != 4 or im0.shape[1] != 3 or im0.shape != im1.shape:
            raise RuntimeError(f"expected two [B,3,H,W] tensors of equal shape, got {tuple(im0.shape)} and {tuple(im1.shape)}")
        if im0.dtype != torch.float32 or im1.dtype != torch.float32:
            raise RuntimeError("ATM-VFI forward expects float32 frames in [0,1]")
        if im0.device != im1.device:
            raise RuntimeError("im0 and im1 are on different devices")

    def forward_normal(self, im0, im1, _ensemble=False):
        """im0, im1: [B,3,H,W] float32 in [0,1] on the model's CUDA device -> the reference's 10-entry dict
        (network_base.py:535-545).  Inference only: outputs carry no autograd graph."""
        self._check_inputs(im0, im1)
        rt = self._runtime
        rt.prepare(self, im0.device, self.precision, self.local_motion_args["window_size"], self.global_motion_args["window_size"])
        B, _, H, W = im0.shape
        with torch.cuda.device(im0.device):
            plan = rt.plan(B, H, W, bool(self.global_motion), _ensemble)
            out = plan.run(im0, im1, use_graph=self.use_cuda_graph)
            return out if self.zero_copy_outputs else clone_outputs(out)

    # ---- fused uint8 path used by demo_2x.inference_2frame ---------------------------------------------
    def interpolate_u8(self, img0, img1, isBGR=True, divisor=64):
        """numpy HxWx3 uint8 pair -> numpy HxWx3 uint8 middle frame.  Same arithmetic as the reference's
        inference_2frame (demo_2x.py:54-87) with the colour flip, /255, replicate padding, un-padding,
        *255 + np.round and the cast done by two small kernels around the forward, so only 3 bytes per
        pixel cross PCIe in each direction.  Pinned staging buffers are cached per shape."""
        import numpy as np
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("inference needs the model on a CUDA device (model.to('cuda')); there is no CPU fallback")
        if img0.shape != img1.shape or img0.ndim != 3 or img0.shape[2] != 3 or img0.dtype != np.uint8:
            raise RuntimeError(f"expected two HxWx3 uint8 frames of equal shape, got {img0.shape} {img0.dtype} and {img1.shape} {img1.dtype}")
        H, W = img0.shape[:2]
        eh, ew = (-H) % divisor, (-W) % divisor
        Hp, Wp, top, left = H + eh, W + ew, eh // 2, ew // 2
        rt = self._runtime
        rt.prepare(self, dev, self.precision, self.local_motion_args["window_size"], self.global_motion_args["window_size"])
        with torch.cuda.device(dev):
            plan = rt.plan(1, Hp, Wp, bool(self.global_motion))
            st = rt.staging(H, W, dev)
            # frames that already live in the pinned staging buffers (pinned_frame_buffers) are uploaded in place; others are
            # copied there first, frame 1 while frame 0 is already on its way to the device
            for src, h, d in ((img0, st["h0"], st["d0"]), (img1, st["h1"], st["d1"])):
                if src.__array_interface__["data"][0] != h.data_ptr():
                    h.numpy()[...] = src
                d.copy_(h, non_blocking=True)
            ops = rt._ops
            ops.u8_to_planar(st["d0"], plan.im0, H, W, Hp, Wp, top, left, isBGR)
            ops.u8_to_planar(st["d1"], plan.im1, H, W, Hp, Wp, top, left, isBGR)
            out = plan.run_inplace(use_graph=self.use_cuda_graph)
            ops.planar_to_u8(out["I_t"], st["dout"], H, W, Hp, Wp, top, left, isBGR)
            st["hout"].copy_(st["dout"], non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return st["hout"].numpy().copy()

    def pinned_frame_buffers(self, H, W):
        """Two HxWx3 uint8 numpy arrays backed by the pinned staging memory of ``interpolate_u8`` / ``inference_2frame``: a decoder
        that writes its frames straight into them saves one host copy per frame (the arrays are reused by the next call)."""
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("pinned frame buffers need the model on a CUDA device")
        st = self._runtime.staging(H, W, dev)
        return st["h0"].numpy(), st["h1"].numpy()

    def interpolate_stream(self, frames, isBGR=True, divisor=64, include_inputs=True):
        """2x interpolation of a frame stream (the loop of demo_2x.py:129-168) as a 2-deep pipeline.

        ``frames``: iterable of HxWx3 uint8 arrays.  Yields uint8 frames in display order: f0, mid(0,1), f1, mid(1,2), ...,
        f_last (only the mids with ``include_inputs=False``).  Per pair the arithmetic is exactly ``inference_2frame``'s.
        Every frame crosses PCIe once: frame k+1 is uploaded from pinned memory on a copy stream while pair (k-1, k) is
        computed, becomes ``im1`` of pair (k, k+1) and is moved to ``im0`` on the device for the next pair; the result
        of a pair is downloaded while the next one runs."""
        import numpy as np
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("inference needs the model on a CUDA device (model.to('cuda')); there is no CPU fallback")
        it = iter(frames)
        try:
            first = np.ascontiguousarray(next(it))
        except StopIteration:
            return
        if first.ndim != 3 or first.shape[2] != 3 or first.dtype != np.uint8:
            raise RuntimeError(f"expected HxWx3 uint8 frames, got {first.shape} {first.dtype}")
        H, W = first.shape[:2]
        eh, ew = (-H) % divisor, (-W) % divisor
        Hp, Wp, top, left = H + eh, W + ew, eh // 2, ew // 2
        rt = self._runtime
        rt.prepare(self, dev, self.precision, self.local_motion_args["window_size"], self.global_motion_args["window_size"])
        with torch.cuda.device(dev):
            reuse = bool(self.stream_encoder_reuse)
            # The generator keeps frame state across yields (the previous frame, and with encoder reuse its features inside the
            # plan's buffers).  With reuse it therefore OWNS a stream plan until it finishes; without reuse it shares the ordinary
            # plan of this shape with forward() / inference_2frame and keeps the previous planar frame in a private buffer, so a
            # consumer that interpolates other pairs of the same shape between two yields cannot disturb the stream.
            plan = rt.acquire_stream_plan(1, Hp, Wp, bool(self.global_motion)) if reuse else rt.plan(1, Hp, Wp, bool(self.global_motion))
            try:
                yield from self._stream_loop(plan, reuse, it, first, H, W, Hp, Wp, top, left, isBGR, include_inputs, dev)
            finally:
                if reuse:
                    rt.release_stream_plan(plan)

    def _stream_loop(self, plan, reuse, it, first, H, W, Hp, Wp, top, left, isBGR, include_inputs, dev):
        import numpy as np
        rt = self._runtime
        ops = rt._ops
        main = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        mk_h = lambda: torch.empty((H, W, 3), dtype=torch.uint8).pin_memory()
        mk_d = lambda: torch.empty((H, W, 3), dtype=torch.uint8, device=dev)
        slots = [dict(h_in=mk_h(), d_in=mk_d(), h_out=mk_h(), d_out=mk_d(), up=torch.cuda.Event(), done=torch.cuda.Event(), down=torch.cuda.Event())
                 for _ in range(2)]
        prev_planar = torch.empty_like(plan.im1)     # frame k as fp32 planar: im0 of pair (k, k+1)

        def upload(slot, frame):                 # host -> pinned -> device, on the copy stream
            s = slots[slot]
            s["up"].synchronize()                # the previous upload from this pinned buffer has finished
            s["h_in"].numpy()[...] = frame
            with torch.cuda.stream(side):
                s["d_in"].copy_(s["h_in"], non_blocking=True)
                s["up"].record(side)

        upload(0, first)
        main.wait_event(slots[0]["up"])
        ops.u8_to_planar(slots[0]["d_in"], plan.im1, H, W, Hp, Wp, top, left, isBGR)      # becomes im0 of the first pair
        prev_planar.copy_(plan.im1)
        if reuse:
            plan.encode_only(use_graph=self.use_cuda_graph)      # features of the first frame; later frames are encoded by the pair step
        slots[0]["done"].record(main)
        prev_frame, pending, k = first, None, 0
        for nxt in it:
            nxt = np.ascontiguousarray(nxt)
            if nxt.shape != first.shape or nxt.dtype != np.uint8:
                raise RuntimeError(f"frame {k + 1} has shape {nxt.shape}, expected {first.shape}")
            if plan is not rt._plans.get(plan.key) and not reuse:
                plan = rt.plan(1, Hp, Wp, bool(self.global_motion))      # the shared plan was evicted / re-packed between two yields
            slot = (k + 1) & 1
            side.wait_event(slots[slot]["done"])     # the pair that used this slot's device buffers has been computed
            upload(slot, nxt)
            s = slots[slot]
            main.wait_event(s["up"])
            plan.im0.copy_(prev_planar)              # frame k was im1 of the previous pair
            ops.u8_to_planar(s["d_in"], plan.im1, H, W, Hp, Wp, top, left, isBGR)
            prev_planar.copy_(plan.im1)
            main.wait_event(s["down"])               # this slot's previous result has left d_out
            out = plan.run_inplace(use_graph=self.use_cuda_graph)
            ops.planar_to_u8(out["I_t"], s["d_out"], H, W, Hp, Wp, top, left, isBGR)
            s["done"].record(main)
            with torch.cuda.stream(side):
                side.wait_event(s["done"])
                s["h_out"].copy_(s["d_out"], non_blocking=True)
                s["down"].record(side)
            if pending is not None:                  # hand out the previous pair while this one runs
                ps, pf = pending
                ps["down"].synchronize()
                if include_inputs:
                    yield pf
                yield ps["h_out"].numpy().copy()
            pending = (s, prev_frame)
            prev_frame = nxt
            k += 1
        if pending is not None:
            ps, pf = pending
            ps["down"].synchronize()
            if include_inputs:
                yield pf
            yield ps["h_out"].numpy().copy()
        if include_inputs:
            yield prev_frame
        main.synchronize()

    # ---- recursive 4x / 8x interpolation on the device (benchmark/davis-vid.py:102-106) -----------------------------------
    def _middle(self, im0, im1):
        """I_t of one pair as a fresh tensor; the other nine outputs stay in the plan's buffers (no clones)."""
        self._check_inputs(im0, im1)
        rt = self._runtime
        rt.prepare(self, im0.device, self.precision, self.local_motion_args["window_size"], self.global_motion_args["window_size"])
        B, _, H, W = im0.shape
        with torch.cuda.device(im0.device):
            plan = rt.plan(B, H, W, bool(self.global_motion), bool(self.ensemble_global_motion))
            return plan.run(im0, im1, use_graph=self.use_cuda_graph)["I_t"].clone()

    def interpolate_recursive(self, im0, im1, levels=2, TTA=False):
        """2^levels x interpolation by recursive 2x, as the reference's DAVIS demo does for 4x (benchmark/davis-vid.py:102-106:
        ``pred025 = model(img0, pred)``, ``pred075 = model(pred, img1)``): im0, im1 [B,3,H,W] float32 in [0,1] on the model's
        device -> the 2^levels - 1 in-between frames in temporal order, all fp32 on the device.  Every level consumes the
        UN-rounded fp32 frames of the level above: nothing goes through uint8 or the host between levels.  ``TTA`` is the
        script's flip augmentation (davis-vid.py:108-112): it replaces the CENTRAL frame by the average with the prediction of
        the flipped pair, after the deeper levels were computed from the plain prediction."""
        if levels < 1:
            raise ValueError("levels must be >= 1")

        def rec(a, b, depth):
            mid = self._middle(a, b)
            if depth == 1:
                return [mid]
            return rec(a, mid, depth - 1) + [mid] + rec(mid, b, depth - 1)

        frames = rec(im0, im1, levels)
        if TTA:
            c = len(frames) // 2
            pf = self._middle(im0.flip(2).flip(3).contiguous(), im1.flip(2).flip(3).contiguous())
            frames[c] = (frames[c] + pf.flip(2).flip(3)) / 2
        return frames

    def interpolate_recursive_u8(self, img0, img1, levels=2, isBGR=True, divisor=64, TTA=False):
        """numpy HxWx3 uint8 pair -> list of 2^levels - 1 numpy uint8 frames (temporal order): ``inference_2frame`` arithmetic at
        both ends (colour flip, /255, replicate padding; *255, np.round, crop), ``interpolate_recursive`` in between.  The pair is
        uploaded once and only finished uint8 frames come back."""
        import numpy as np
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("inference needs the model on a CUDA device (model.to('cuda')); there is no CPU fallback")
        if img0.shape != img1.shape or img0.ndim != 3 or img0.shape[2] != 3 or img0.dtype != np.uint8:
            raise RuntimeError(f"expected two HxWx3 uint8 frames of equal shape, got {img0.shape} {img0.dtype} and {img1.shape} {img1.dtype}")
        H, W = img0.shape[:2]
        eh, ew = (-H) % divisor, (-W) % divisor
        Hp, Wp, top, left = H + eh, W + ew, eh // 2, ew // 2
        rt = self._runtime
        rt.prepare(self, dev, self.precision, self.local_motion_args["window_size"], self.global_motion_args["window_size"])
        with torch.cuda.device(dev):
            ops = rt._ops
            st = rt.staging(H, W, dev)
            a, b = torch.empty(1, 3, Hp, Wp, device=dev), torch.empty(1, 3, Hp, Wp, device=dev)
            for src, h, d, dst in ((img0, st["h0"], st["d0"], a), (img1, st["h1"], st["d1"], b)):
                if src.__array_interface__["data"][0] != h.data_ptr():
                    h.numpy()[...] = src
                d.copy_(h, non_blocking=True)
                ops.u8_to_planar(d, dst, H, W, Hp, Wp, top, left, isBGR)
            frames = self.interpolate_recursive(a, b, levels, TTA)
            out_d = torch.empty((len(frames), H, W, 3), dtype=torch.uint8, device=dev)
            for i, f in enumerate(frames):
                ops.planar_to_u8(f.contiguous(), out_d[i], H, W, Hp, Wp, top, left, isBGR)
            out_h = out_d.cpu().numpy()
            return [out_h[i] for i in range(len(frames))]

    def invalidate(self):
        """Drop the packed weights / plans / CUDA graphs (needed only after ``p.data`` edits that bypass autograd's version counter)."""
        self._runtime.invalidate()

    def forward_global_ensemble(self, im0, im1):
        """forward with the multi-scale global-motion ensemble (network_base.py:564-712): the global flows are estimated at
        input scales 1, 1/2 and 1/4 and, per sample, the scale that aligns the two frames best is kept (selected on the
        device).  H and W must be multiples of 64.  ``im_t_list`` and the warped lists hold 4 scales, as in the reference."""
        return self.forward_normal(im0, im1, _ensemble=True)
