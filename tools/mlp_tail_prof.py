"""Time atmvfi_mlp_tail on the Base 1080p token grids and print the per-role cycle counters of CTA 0 (ATMVFI_MT_PROF=1).
usage: ATMVFI_MT_PROF=1 python tools/mlp_tail_prof.py [tf32|f16]"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'atm-vfi_b200'), os.path.join(ROOT, 'atm-vfi_b200', 'network'), os.path.join(ROOT, 'tests')]
import torch
from atmvfi import _lib, pack
from atmvfi.ops import CudaOps, Map, PackedGemm
prec = sys.argv[1] if len(sys.argv) > 1 else 'tf32'
f16 = prec == 'f16'
dev = torch.device('cuda:0')
cu = CudaOps(dev, _lib.F16 if f16 else _lib.TF32)
NAMES = ["prod wait rawEmpty", "prod wait bEmpty", "mma wait tEmpty", "mma wait aFull", "mma wait bFull", "mma issue", "conv wait rawFull",
         "conv wait aEmpty", "conv compute", "epi wait tFull", "epi work"]
for (B, H, W, C, hid) in ((2, 136, 240, 384, 1536), (2, 68, 120, 672, 2688)):
    g = torch.Generator().manual_seed(1)
    P = {"fc2.weight": torch.randn(C, hid, generator=g) / hid ** 0.5, "fc2.bias": torch.randn(C, generator=g) * 0.1,
         "dw.weight": torch.randn(hid, 1, 3, 3, generator=g) / 3.0, "dw.bias": torch.randn(hid, generator=g) * 0.1}
    fc2 = pack.pack_linear(P, ["fc2"]); dw_w, dw_b = pack.pack_dw(P, "dw")
    c = lambda t: None if t is None else t.cuda()
    fc2 = PackedGemm(fc2.name, fc2.ksize, fc2.split, fc2.Cout, fc2.shuffle, c(fc2.w32), c(fc2.bias), c(fc2.prelu))
    dt = torch.float16 if f16 else torch.float32
    h = Map(torch.randn(B, H, W, hid, device=dev).to(dt), 0, hid); x = Map(torch.randn(B, H, W, C, device=dev).to(dt), 0, C)
    out = cu.new_map(B, H, W, C)
    for _ in range(3): cu.mlp_tail(h, dw_w.cuda(), dw_b.cuda(), fc2, x, out)
    torch.cuda.synchronize()
    buf = (ctypes.c_uint64 * 336)()
    cu.lib.atmvfi_mlp_tail_prof_read(buf)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    R = 10
    e0.record()
    for _ in range(R): cu.mlp_tail(h, dw_w.cuda(), dw_b.cuda(), fc2, x, out)
    e1.record(); torch.cuda.synchronize()
    print(f"{prec} {B}x{H}x{W} C={C} hid={hid}: {e0.elapsed_time(e1) / R * 1000:.1f} us per launch")
    if cu.lib.atmvfi_mlp_tail_prof_read(buf) == 0:
        for i, n in enumerate(NAMES): print(f"   {n:22s} {buf[i] / R / 1000:10.1f} kcycles per launch")
        st = [buf[16 + 2 * i] for i in range(148)]; en = [buf[17 + 2 * i] for i in range(148)]
        t0 = min(st)
        print('   CTA start offsets (us):', sorted(set(round((v - t0) / 1000) for v in st)))
        print('   CTA durations (us): min %.1f max %.1f; last end %.1f' % (min(e - b for b, e in zip(st, en)) / 1000, max(e - b for b, e in zip(st, en)) / 1000, (max(en) - t0) / 1000))
