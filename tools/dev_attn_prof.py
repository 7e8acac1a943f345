import sys
sys.path[:0] = ['atm-vfi_b200', 'atm-vfi_b200/network', 'oracle', 'tests']
import torch
from atmvfi import _lib
from atmvfi.ops import CudaOps, Map, WinGeom
cu = CudaOps(torch.device('cuda:0'), _lib.TF32)
hd, heads = 48, 8; C = hd * heads
geo = WinGeom(2, 136, 240, 8, 4)
qkv = Map(torch.randn(1, 1, geo.rows, 3 * C, device='cuda')); out = Map(torch.zeros(1, 1, geo.rows, C, device='cuda'))
mo = Map(torch.zeros(1, 136, 240, 8, device='cuda')); scratch = torch.empty(geo.rows * heads * 2, device='cuda')
mix = [torch.randn(4, 8, device='cuda'), torch.randn(4, device='cuda'), torch.randn(4, device='cuda'), torch.randn(1, device='cuda')]
rc = torch.zeros(2, 64, 64, device='cuda')
for _ in range(3):
    cu.window_attention(qkv, out, geo, heads, True, rc, mix, mo, 0, scratch, rc_closed_form=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): cu.window_attention(qkv, out, geo, heads, True, rc, mix, mo, 0, scratch, rc_closed_form=True)
e1.record(); torch.cuda.synchronize(); print('ms', e0.elapsed_time(e1) / 10)
