"""nn.Module shells that own the reference-named parameters.

The reference builds its networks from nn.Conv2d / nn.Linear / nn.PReLU sub-modules and runs them
eagerly.  Here modules only OWN parameters (so ``state_dict`` / ``load_state_dict(strict=True)`` /
``.to()`` / ``requires_grad_`` behave as callers of the reference expect); the arithmetic is
done by the CUDA engine.  ``ParamTree`` builds the nested module tree straight from the flat schema of
arch.param_schema().
"""
from __future__ import annotations

import math
from typing import Dict, Iterable, Tuple

import torch
import torch.nn as nn


def relative_coord_buffer(ws: int) -> torch.Tensor:
    """[1,1,2,N,N]: (key - query) x- and y-offsets inside a ws x ws window, row-major tokens
    (what attention.py:150-165 registers as a persistent buffer)."""
    idx = torch.arange(ws * ws)
    px, py = (idx % ws).float(), (idx // ws).float()
    return torch.stack([px[None, :] - px[:, None], py[None, :] - py[:, None]], 0)[None, None].contiguous()


def _init_tensor(shape: Tuple[int, ...], kind: str, owner_shape=None) -> torch.Tensor:
    """Same init families as the reference (default torch init for the CNN stacks, attention.py:101-114
    for everything that goes through ``_init_weights``)."""
    t = torch.empty(shape)
    if kind == "conv":
        nn.init.kaiming_uniform_(t, a=math.sqrt(5))
    elif kind == "deconv":
        nn.init.kaiming_uniform_(t, a=math.sqrt(5))
    elif kind in ("conv_tf", "dw"):
        fan_out = shape[2] * shape[3] * shape[0]
        if kind == "dw":
            fan_out //= shape[0]
        t.normal_(0, math.sqrt(2.0 / fan_out))
    elif kind == "linear":
        nn.init.trunc_normal_(t, std=0.02)
    elif kind == "bias":
        fan_in = owner_shape[1] * owner_shape[2] * owner_shape[3]
        bound = 1 / math.sqrt(fan_in) if fan_in > 0 else 0
        t.uniform_(-bound, bound)
    elif kind in ("bias_tf", "ln_b"):
        t.zero_()
    elif kind == "ln_w":
        t.fill_(1.0)
    elif kind == "prelu":
        t.fill_(0.25)
    else:
        raise ValueError(kind)
    return t


# Bumped whenever a tensor OBJECT of any ParamTree may have been replaced (attribute assignment, .to() / .cuda() / .half(),
# load_state_dict).  runtime.Runtime caches the list of a network's 236 tensors and re-walks the module tree only when this moves;
# in-place value changes are caught per call through (data_ptr, _version).
STRUCT_EPOCH = [0]


class ParamTree(nn.Module):
    """A module whose children / parameters are created from dotted names."""

    def __init__(self):
        super().__init__()

    def __setattr__(self, name, value):
        if isinstance(value, torch.Tensor):
            STRUCT_EPOCH[0] += 1
        super().__setattr__(name, value)

    def register_parameter(self, name, param):
        STRUCT_EPOCH[0] += 1
        super().register_parameter(name, param)

    def register_buffer(self, name, tensor, persistent=True):
        STRUCT_EPOCH[0] += 1
        super().register_buffer(name, tensor, persistent)

    def _apply(self, fn, recurse=True):
        STRUCT_EPOCH[0] += 1
        return super()._apply(fn, recurse)

    def load_state_dict(self, *a, **k):
        STRUCT_EPOCH[0] += 1
        return super().load_state_dict(*a, **k)

    def _descend(self, path: Iterable[str]) -> "ParamTree":
        node = self
        for part in path:
            if part not in node._modules:
                node.add_module(part, ParamTree())
            node = node._modules[part]
        return node

    def populate(self, schema: Dict[str, tuple], prefix: str = "") -> None:
        for name, (shape, kind) in schema.items():
            if prefix:
                if not name.startswith(prefix + "."):
                    continue
                rel = name[len(prefix) + 1:]
            else:
                rel = name
            *path, leaf = rel.split(".")
            node = self._descend(path)
            if kind == "buffer":
                ws = int(round(math.sqrt(shape[-1])))
                node.register_buffer(leaf, relative_coord_buffer(ws))
            else:
                owner = schema.get(name[: -len("bias")] + "weight", (None,))[0] if leaf == "bias" else None
                node.register_parameter(leaf, nn.Parameter(_init_tensor(tuple(shape), kind, owner)))

    def forward(self, *a, **k):
        raise RuntimeError("parameter container: the forward pass runs in the CUDA engine (Network.forward)")
