"""Times atmvfi_dwconv3x3_gelu alone (Base 1080p local-branch shape).  usage: ATMVFI_DW_V=4 ATMVFI_DW_DEPTH=2 python tools/bench_dwconv.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'atm-vfi_b200')]
import torch
from atmvfi import _lib
from atmvfi.ops import CudaOps, Map
ops = CudaOps(torch.device('cuda:0'), _lib.TF32)
for (B, H, W, C) in ((2, 136, 240, 1536), (2, 68, 120, 2688)):
    x = Map(torch.randn(B, H, W, C, device='cuda'))
    o = Map(torch.empty(B, H, W, C, device='cuda'))
    w, b = torch.randn(9, C, device='cuda') * 0.3, torch.randn(C, device='cuda')
    for _ in range(5): ops.dwconv_gelu(x, o, w, b)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): ops.dwconv_gelu(x, o, w, b)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 50 * 1e3
    print(f"V={os.environ.get('ATMVFI_DW_V','4')} depth={os.environ.get('ATMVFI_DW_DEPTH','2')} {B}x{H}x{W}x{C}: {us:.1f} us  {2*4*B*H*W*C/us/1e3:.0f} GB/s")
