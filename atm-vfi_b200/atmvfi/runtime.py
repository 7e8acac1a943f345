"""Per-module engine cache: packs weights lazily, re-packs when parameters change, caches plans per shape."""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from . import _lib
from .arch import Arch
from .engine import PackedModel, Plan, clone_outputs
from .modules import STRUCT_EPOCH
from .ops import CudaOps

PRECISIONS = {"fp32": _lib.FP32, "tf32": _lib.TF32, "fp32x3": _lib.TF32X3, "f16": _lib.F16}


class Runtime:
    def __init__(self, arch: Arch):
        self.arch = arch
        self._sig = None
        self._model: Optional[PackedModel] = None
        self._ops: Optional[CudaOps] = None
        self._plans: Dict[Tuple, Plan] = {}
        self._stream_plans: Dict[Tuple, list] = {}
        self._staging: Dict[Tuple, dict] = {}
        self._tensors = None
        self._epoch = -1

    def invalidate(self) -> None:
        """Forget the packed weights, plans and CUDA graphs; the next forward re-packs from the module's current tensors.
        Needed only after a weight update that bypasses autograd's version counter (``p.data.copy_()`` / ``p.data.mul_()``,
        EMA-style updates): ordinary in-place updates, ``load_state_dict`` and ``.to()`` are detected automatically."""
        self._sig, self._tensors = None, None
        self._plans.clear()
        self._stream_plans.clear()

    def _signature(self, module: torch.nn.Module, extra) -> Tuple:
        # the module tree (236 tensors) is walked only when a tensor object may have been replaced (modules.STRUCT_EPOCH);
        # per call the cached tensors are fingerprinted by address and in-place version: ~0.08 ms instead of 1.2 ms
        if self._tensors is None or self._epoch != STRUCT_EPOCH[0]:
            self._tensors = [t for _, t in list(module.named_parameters()) + list(module.named_buffers())]
            self._epoch = STRUCT_EPOCH[0]
        return (tuple((t.data_ptr(), t._version) for t in self._tensors), extra)

    def prepare(self, module: torch.nn.Module, device: torch.device, precision: str, local_ws: int, global_ws: int) -> None:
        if precision not in PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(PRECISIONS)}, got {precision!r}")
        sig = self._signature(module, (str(device), precision, local_ws, global_ws))
        if sig == self._sig:
            return
        self._ops = CudaOps(device, PRECISIONS[precision])          # raises without CUDA / the built library
        sd = {k: v.detach() for k, v in module.state_dict().items()}
        for k, v in sd.items():
            if v.device != device:
                raise RuntimeError(f"parameter {k} is on {v.device} but the inputs are on {device}")
        self._model = PackedModel(self.arch, sd, local_ws, global_ws, with_global=True)
        self._plans.clear()
        self._stream_plans.clear()
        self._sig = sig

    def acquire_stream_plan(self, B: int, H: int, W: int, global_motion: bool) -> Plan:
        """A video-stream plan (encoder features persist between its steps) that no other stream is using.  Every
        ``interpolate_stream`` generator owns one until it finishes, so two streams of one shape - or a stream interleaved with
        ``inference_2frame`` calls - never see each other's frames or features."""
        key = (B, H, W, bool(global_motion))
        pool = self._stream_plans.setdefault(key, [])
        for p in pool:
            if not p.in_use:
                p.in_use = True
                return p
        try:
            p = Plan(self._ops, self._model, B, H, W, bool(global_motion), False, True)
        except torch.cuda.OutOfMemoryError:
            self._plans.clear()
            for k in list(self._stream_plans):
                self._stream_plans[k] = [q for q in self._stream_plans[k] if q.in_use]
            torch.cuda.empty_cache()
            p = Plan(self._ops, self._model, B, H, W, bool(global_motion), False, True)
            pool = self._stream_plans.setdefault(key, [])
        p.in_use = True
        pool.append(p)
        return p

    @staticmethod
    def release_stream_plan(p: Plan) -> None:
        p.in_use = False

    def plan(self, B: int, H: int, W: int, global_motion: bool, ensemble: bool = False, stream: bool = False) -> Plan:
        key = (B, H, W, bool(global_motion), bool(ensemble and global_motion), bool(stream))
        p = self._plans.get(key)
        if p is None:
            if len(self._plans) >= 4:            # plans own all activation buffers; keep the cache small
                self._plans.pop(next(iter(self._plans)))
            try:
                p = Plan(self._ops, self._model, B, H, W, bool(global_motion), bool(ensemble), bool(stream))
            except torch.cuda.OutOfMemoryError:
                # plans own all their activation buffers (78 GB for Base at 4K): make room by dropping the cached ones, once
                if not self._plans:
                    raise
                self._plans.clear()
                torch.cuda.empty_cache()
                p = Plan(self._ops, self._model, B, H, W, bool(global_motion), bool(ensemble), bool(stream))
            self._plans[key] = p
        return p

    def staging(self, H: int, W: int, device: torch.device) -> dict:
        """Pinned host + device uint8 frame buffers for the fused uint8 path."""
        key = (H, W, str(device))
        st = self._staging.get(key)
        if st is None:
            if len(self._staging) >= 4:
                self._staging.pop(next(iter(self._staging)))
            mk_h = lambda: torch.empty((H, W, 3), dtype=torch.uint8).pin_memory()
            mk_d = lambda: torch.empty((H, W, 3), dtype=torch.uint8, device=device)
            st = self._staging[key] = dict(h0=mk_h(), h1=mk_h(), hout=mk_h(), d0=mk_d(), d1=mk_d(), dout=mk_d())
        return st
