"""GPU debug: tcgen05 gemm_conv vs the fp32 SIMT kernel on the same inputs."""
import sys
sys.path[:0] = ['atm-vfi_b200', 'atm-vfi_b200/network', 'oracle', 'tests']
import torch
from atmvfi import _lib, pack
from atmvfi.ops import CudaOps, Map, WinGeom, PackedGemm
from gpu_util import rand_map, to_gpu, max_err

dev = torch.device('cuda:0')
f32, tc = CudaOps(dev, _lib.FP32), CudaOps(dev, _lib.TF32)
g = torch.Generator().manual_seed(0)

def conv_case(B, H, W, split, Co, k, stride, dil, shuffle=False):
    ci = sum(split)
    if shuffle:
        P = {"d.0.weight": torch.randn(ci, Co, 2, 2, generator=g) / ci ** 0.5, "d.0.bias": torch.randn(Co, generator=g) * 0.1, "d.1.weight": torch.rand(Co, generator=g) * 0.5}
        w = pack.pack_deconvp(P, "d", split=split)
    else:
        P = {"c.weight": torch.randn(Co, ci, k, k, generator=g) / (ci * k * k) ** 0.5, "c.bias": torch.randn(Co, generator=g) * 0.1, "p": torch.rand(Co, generator=g) * 0.5}
        w = pack.pack_conv(P, "c", split=split, prelu="p")
    wg = PackedGemm(w.name, w.ksize, w.split, w.Cout, w.shuffle, w.w32.cuda(), w.bias.cuda(), w.prelu.cuda())
    srcs = [to_gpu(rand_map(B, H, W, c, gen=g)) for c in split]
    pad = dil * (k - 1) // 2
    Ho = (H + 2 * pad - dil * (k - 1) - 1) // stride + 1
    Wo = (W + 2 * pad - dil * (k - 1) - 1) // stride + 1
    if shuffle: Ho, Wo = 2 * H, 2 * W
    o1 = to_gpu(Map(torch.zeros(B, Ho, Wo, (Co + 3) // 4 * 4), 0, Co)); o2 = to_gpu(Map(torch.zeros(B, Ho, Wo, (Co + 3) // 4 * 4), 0, Co))
    f32.gemm_conv(srcs, wg, o1, stride=stride, dil=dil)
    tc.gemm_conv(srcs, wg, o2, stride=stride, dil=dil)
    torch.cuda.synchronize()
    e = max_err(o1, o2); ref = o1.view().abs().max().item()
    bad = (o1.view() - o2.view()).abs() > 0.02 * max(ref, 1)
    print(f"B{B} {H}x{W} split{split} Co{Co} k{k} s{stride} d{dil} shuffle={shuffle}: max err {e:.3e} (ref max {ref:.2f}) bad {bad.float().mean().item():.4f}", flush=True)
    if bad.any():
        idx = bad.nonzero()[:5].tolist(); print("   first bad idx", idx)
    return e

cases = [
    (1, 8, 16, [32], 32, 1, 1, 1), (1, 8, 16, [32], 32, 3, 1, 1), (1, 16, 32, [64], 64, 3, 1, 1), (2, 17, 23, [40], 24, 3, 1, 1),
    (1, 9, 13, [8, 20, 20], 36, 3, 1, 1), (1, 16, 16, [96, 48, 48, 192], 384, 1, 1, 1), (1, 20, 28, [101, 15], 64, 3, 1, 1),
    (1, 1, 1000, [384], 1152, 1, 1, 1), (1, 64, 64, [64], 5, 1, 1, 1), (1, 40, 30, [389], 389, 3, 1, 1),
    (1, 32, 40, [24], 48, 3, 2, 1), (2, 32, 48, [48], 48, 3, 4, 1), (2, 32, 48, [48], 48, 3, 4, 2),
]
which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "conv"):
    for c in cases: conv_case(*c)
if which in ("all", "deconv"):
    conv_case(2, 7, 9, [37], 21, 1, 1, 1, True); conv_case(1, 17, 30, [384, 384, 5], 389, 1, 1, 1, True)
if which in ("all", "speed"):
    for (H, W, ci, co, k) in ((136, 240, 576, 576, 3), (1088, 1920, 101, 101, 3), (272, 480, 389, 389, 3), (1, 65280, 384, 1536, 1)):
        P = {"c.weight": torch.randn(co, ci, k, k, generator=g) / (ci * k * k) ** 0.5, "c.bias": torch.randn(co, generator=g) * 0.1, "p": torch.rand(co, generator=g)}
        w = pack.pack_conv(P, "c", prelu="p"); wg = PackedGemm(w.name, w.ksize, w.split, w.Cout, False, w.w32.cuda(), w.bias.cuda(), w.prelu.cuda())
        src = Map(torch.randn(1, H, W, (ci + 3) // 4 * 4, device=dev), 0, ci); out = Map(torch.zeros(1, H, W, (co + 3) // 4 * 4, device=dev), 0, co)
        tc.recording = rec = []; tc.gemm_conv([src], wg, out); tc.recording = None
        for _ in range(3): tc.replay(rec)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); n = 10
        for _ in range(n): tc.replay(rec)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n; fl = 2.0 * H * W * ci * co * k * k
        print(f"speed {H}x{W} {ci}->{co} k{k}: {ms:.3f} ms  {fl/ms/1e9:.1f} TFLOP/s", flush=True)
if which == "speed2":
    import time
    shapes = [(1, 65280, 384, 1536, 1), (1, 65280, 1536, 384, 1), (1, 65280, 384, 1152, 1), (1, 65280, 384, 384, 1), (544, 960, 197, 197, 3), (1088, 1920, 101, 101, 3)]
    recs = []
    for (H, W, ci, co, k) in shapes:
        P = {"c.weight": torch.randn(co, ci, k, k, generator=g) / (ci * k * k) ** 0.5, "c.bias": torch.randn(co, generator=g) * 0.1, "p": torch.rand(co, generator=g)}
        w = pack.pack_conv(P, "c", prelu="p"); wg = PackedGemm(w.name, w.ksize, w.split, w.Cout, False, w.w32.cuda(), w.bias.cuda(), w.prelu.cuda())
        src = Map(torch.randn(1, H, W, (ci + 3) // 4 * 4, device=dev), 0, ci); out = Map(torch.zeros(1, H, W, (co + 3) // 4 * 4, device=dev), 0, co)
        tc.recording = rec = []; tc.gemm_conv([src], wg, out); tc.recording = None
        recs.append((rec, 2.0 * H * W * ci * co * k * k, (H, W, ci, co, k)))
    t0 = time.time()
    while time.time() - t0 < 1.5:
        for rec, _, _ in recs: tc.replay(rec)
    torch.cuda.synchronize()
    for rec, fl, shp in recs:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); n = 20
        for _ in range(n): tc.replay(rec)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        print(f"speed2 {shp}: {ms*1e3:.1f} us  {fl/ms/1e9:.1f} TFLOP/s", flush=True)
