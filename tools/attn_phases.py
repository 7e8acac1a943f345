"""Phase breakdown of the tcgen05 window-attention kernel (needs ATMVFI_ATTN_PROF=1).  usage: ATMVFI_ATTN_PROF=1 python tools/attn_phases.py"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'atm-vfi_b200')]
import torch
from atmvfi import _lib
from atmvfi.ops import CudaOps, Map, WinGeom
ops = CudaOps(torch.device('cuda:0'), _lib.TF32)
names = ["stage Q/K/V", "S=QK^T (issue+wait)", "softmax max", "softmax exp + P", "O=PV (issue+wait)", "output"]
for (B2, H, W, ws, shift, Cc, cross) in ((2, 136, 240, 8, 4, 384, True), (2, 68, 120, 12, 6, 672, True)):
    g = WinGeom(B2, H, W, ws, shift)
    qkv = Map(torch.randn(1, 1, g.rows, 3 * Cc, device='cuda'))
    out = Map(torch.empty(1, 1, g.rows, Cc, device='cuda'))
    hm = os.environ.get('HEAD_MAJOR', '1') != '0'
    for _ in range(3): ops.window_attention(qkv, out, g, 8, cross, head_major=hm)
    torch.cuda.synchronize()
    buf = (C.c_uint64 * 6)()
    ops.lib.atmvfi_attn_prof_read(buf)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): ops.window_attention(qkv, out, g, 8, cross, head_major=hm)
    e1.record(); torch.cuda.synchronize()
    rc = ops.lib.atmvfi_attn_prof_read(buf)
    N = ws * ws
    nwin = g.rows // N
    ctas = ((nwin + (128 // N if N <= 64 else 1) - 1) // (128 // N if N <= 64 else 1)) * 8 * (1 if N <= 128 else (N + 127) // 128)
    print(f"ws={ws} C={Cc} head_major={hm}: {e0.elapsed_time(e1) / 10 * 1e3:.0f} us per launch, {ctas} CTAs; cycles per CTA by phase (rc={rc}):")
    tot = sum(buf)
    for n, v in zip(names, buf):
        print(f"   {n:24s} {v / 10 / ctas:9.0f}  {100 * v / max(tot, 1):5.1f} %")
