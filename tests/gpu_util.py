"""Helpers for the -m gpu parity tests: run one operator call on the CUDA library and on the CPU contract
emulation with mirrored buffers."""
import torch

from atmvfi.ops import Map


def to_gpu(x):
    if isinstance(x, Map):
        return Map(x.t.cuda(), x.c0, x.C)
    if isinstance(x, torch.Tensor):
        return x.cuda()
    if isinstance(x, (list, tuple)):
        return type(x)(to_gpu(v) for v in x)
    return x


def rand_map(B, H, W, C, pitch=None, gen=None, scale=1.0):
    pitch = pitch or (C + 3) // 4 * 4
    t = torch.randn(B, H, W, pitch, generator=gen) * scale
    return Map(t, 0, C)


def max_err(a, b):
    a = a.view() if isinstance(a, Map) else a
    b = b.view() if isinstance(b, Map) else b
    return (a.detach().cpu().float() - b.detach().cpu().float()).abs().max().item()
