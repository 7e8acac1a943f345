"""precision "f16": fp16 STORAGE of the channels-last feature maps + tcgen05 kind::f16 MMAs (fp32 accumulation); flows, masks,
images, q|k|v and the 5-channel motion heads stay fp32.  fp16 carries the same 10-bit mantissa as TF32, so the mode is held to
the tf32 tolerances end to end; per operator the reference is the fp32 contract emulation evaluated on the SAME fp16-valued
inputs (and fp16-rounded weights), and the only allowed difference is the final rounding of a stored value to fp16."""
import glob
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import atmvfi_oracle as oracle
import weights
from atmvfi import _lib, pack
from atmvfi.ops import CudaOps, Map, WinGeom, PackedGemm
from emul_ops import EmulOps
from gpu_util import to_gpu
from test_gpu_forward import CASES, TOL, _net, psnr
from test_gpu_ops_tc import TC_CONV_CASES, _conv_weights, _pg_to_gpu, gen, tol_k

ULP = 2.0 ** -10          # one fp16 rounding step relative to the value (round-to-nearest: half of it, the bound leaves slack)


@pytest.fixture(scope="module")
def ops():
    return CudaOps(torch.device("cuda:0"), _lib.F16), EmulOps()


def h_map(B, H, W, C, g, pitch=None, scale=1.0):
    """fp16 map (pitch a multiple of 8 halves) and its fp32 twin holding the same values."""
    pitch = pitch or (C + 7) // 8 * 8
    t = (torch.randn(B, H, W, pitch, generator=g) * scale).half()
    return Map(t, 0, C), Map(t.float(), 0, C)


def f32_out_like(B, H, W, C):
    return Map(torch.zeros(B, H, W, (C + 3) // 4 * 4), 0, C)


def half_weights(w: PackedGemm) -> PackedGemm:
    """The layer as the fp16 kernel sees it: weights rounded to fp16 (bias / slopes stay fp32)."""
    return PackedGemm(w.name, w.ksize, w.split, w.Cout, w.shuffle, w.w32.half().float(), w.bias, w.prelu)


def close(got: torch.Tensor, ref: torch.Tensor, k: int, stored_half: bool):
    err = (got.float().cpu() - ref).abs()
    bound = tol_k(k) + (ULP * ref.abs() if stored_half else 0)
    assert bool((err <= bound).all()), float((err - bound).max())


@pytest.mark.parametrize("B,H,W,split,Co,k,stride,dil", TC_CONV_CASES)
def test_f16_conv(ops, B, H, W, split, Co, k, stride, dil):
    cu, em = ops
    g = gen(31)
    w = pack.pack_conv(_conv_weights(sum(split), Co, k, g), "c", split=split, prelu="p")
    pairs = [h_map(B, H, W, c, g, pitch=(c + 7) // 8 * 8 + 8 * (i % 2)) for i, c in enumerate(split)]
    pad = dil * (k - 1) // 2
    Ho = (H + 2 * pad - dil * (k - 1) - 1) // stride + 1
    Wo = (W + 2 * pad - dil * (k - 1) - 1) // stride + 1
    ref = f32_out_like(B, Ho, Wo, Co)
    em.gemm_conv([p[1] for p in pairs], half_weights(w), ref, stride=stride, dil=dil)
    for out_f32 in (False, True):
        out = Map(torch.zeros(B, Ho, Wo, (Co + 7) // 8 * 8, dtype=torch.float32 if out_f32 else torch.float16).cuda(), 0, Co)
        rec = cu.recording = []
        cu.gemm_conv(to_gpu([p[0] for p in pairs]), _pg_to_gpu(w), out, stride=stride, dil=dil, out_f32=out_f32)
        cu.recording = None
        assert rec[0][3][0].precision == _lib.F16
        cu.round_outputs = False
        cu.replay(rec)
        cu.round_outputs = None
        close(out.view(), ref.view(), k * k * sum(split), stored_half=not out_f32)


def test_f16_transposed_dual_residual_window_head32(ops):
    cu, em = ops
    g = gen(32)
    for split, Co, H, W in (([37], 21, 7, 9), ([384, 384, 5], 37, 7, 9), ([389], 197, 34, 60)):
        ci = sum(split)
        P = {"d.0.weight": torch.randn(ci, Co, 2, 2, generator=g) / ci ** 0.5, "d.0.bias": torch.randn(Co, generator=g) * 0.1,
             "d.1.weight": torch.rand(Co, generator=g) * 0.5}
        w = pack.pack_deconvp(P, "d", split=split)
        pairs = [h_map(2, H, W, c, g) for c in split]
        ref = f32_out_like(2, 2 * H, 2 * W, Co)
        em.gemm_conv([p[1] for p in pairs], half_weights(w), ref)
        out = Map(torch.zeros(2, 2 * H, 2 * W, (Co + 7) // 8 * 8, dtype=torch.float16).cuda(), 0, Co)
        cu.gemm_conv(to_gpu([p[0] for p in pairs]), _pg_to_gpu(w), out)
        close(out.view(), ref.view(), ci, True)
    # decoder level: raw fp16 output + PReLU'd fp16 copy + fp32 copy of the last 5 channels (flows + occlusion logit)
    ci, co, H, W = 197, 101, 40, 56
    w = pack.pack_conv(_conv_weights(ci, co, 3, g), "c")
    (xh, xf), slopes = h_map(2, H, W, ci, g), torch.rand(co, generator=g)
    r1, r2 = f32_out_like(2, H, W, co), f32_out_like(2, H, W, co)
    em.gemm_conv([xf], half_weights(w), r1, act=False, out2=r2, prelu2=slopes)
    o1 = Map(torch.zeros(2, H, W, 104, dtype=torch.float16).cuda(), 0, co)
    o2 = Map(torch.zeros(2, H, W, 104, dtype=torch.float16).cuda(), 0, co)
    hd = Map(torch.zeros(2, H, W, 8).cuda(), 0, 5)
    cu.gemm_conv([to_gpu(xh)], _pg_to_gpu(w), o1, act=False, out2=o2, prelu2=slopes.cuda(), head32=hd, head32_c0=co - 5)
    close(o1.view(), r1.view(), 9 * ci, True)
    close(o2.view(), r2.view(), 9 * ci, True)
    close(hd.view(), r1.view()[..., co - 5:], 9 * ci, False)             # the fp32 copy is NOT rounded to fp16
    # linear + residual, and + window reverse
    for (ci, co, rows) in ((40, 52, 333), (384, 384, 4100), (1536, 384, 700)):
        Pl = {"l.weight": torch.randn(co, ci, generator=g) / ci ** 0.5, "l.bias": torch.randn(co, generator=g)}
        wl = pack.pack_linear(Pl, ["l"])
        (xh, xf), (rh, rf) = h_map(1, 1, rows, ci, g), h_map(1, 1, rows, co, g)
        ref = f32_out_like(1, 1, rows, co)
        em.gemm_conv([xf], half_weights(wl), ref, act=False, residual=rf)
        out = Map(torch.zeros(1, 1, rows, (co + 7) // 8 * 8, dtype=torch.float16).cuda(), 0, co)
        cu.gemm_conv([to_gpu(xh)], _pg_to_gpu(wl), out, act=False, residual=to_gpu(rh))
        close(out.view(), ref.view(), ci, True)
    geo = WinGeom(2, 68, 120, 12, 6)
    Pl = {"l.weight": torch.randn(96, 96, generator=g) * 0.1, "l.bias": torch.randn(96, generator=g)}
    wl = pack.pack_linear(Pl, ["l"])
    (xh, xf), (rh, rf) = h_map(1, 1, geo.rows, 96, g), h_map(1, 1, geo.rows, 96, g)
    ref = f32_out_like(2, 68, 120, 96)
    em.gemm_conv([xf], half_weights(wl), ref, act=False, residual=rf, win=geo)
    out = Map(torch.zeros(2, 68, 120, 96, dtype=torch.float16).cuda(), 0, 96)
    cu.gemm_conv([to_gpu(xh)], _pg_to_gpu(wl), out, act=False, residual=to_gpu(rh), win=geo)
    close(out.view(), ref.view(), 96, True)
    # head-major q | k | v (fp32 output)
    heads, hd_, rows = 8, 48, 1000
    C = heads * hd_
    Pl = {"q.weight": torch.randn(C, C, generator=g) / C ** 0.5, "kv.weight": torch.randn(2 * C, C, generator=g) / C ** 0.5}
    wl = pack.pack_linear(Pl, ["q", "kv"], bias=False)
    xh, xf = h_map(1, 1, rows, C, g)
    ref = Map(torch.zeros(1, 1, rows, 3 * C), 0, 3 * C)
    em.gemm_conv([xf], half_weights(wl), ref, act=False, qkv_heads=heads)
    out = Map(torch.zeros(1, 1, rows, 3 * C).cuda(), 0, 3 * C)
    cu.round_outputs = False
    cu.gemm_conv([to_gpu(xh)], _pg_to_gpu(wl), out, act=False, qkv_heads=heads, out_f32=True)
    cu.round_outputs = None
    close(out.t, ref.t, C, False)


def test_f16_streaming_kernels(ops):
    """LayerNorm (+ window gather), depth-wise 3x3 + GELU, the NHWC token warp, the first conv, the image packer, the cast: fp16 in /
    out against the fp32 contract on the same values."""
    cu, em = ops
    g = gen(33)
    for C in (224, 384, 672):
        (xh, xf), ref = h_map(1, 3, 50, C, g, scale=3.0), f32_out_like(1, 3, 50, C)
        gamma, beta = torch.randn(C, generator=g), torch.randn(C, generator=g)
        em.layernorm(xf, ref, gamma, beta)
        out = Map(torch.zeros(1, 3, 50, C, dtype=torch.float16).cuda(), 0, C)
        cu.layernorm(to_gpu(xh), out, gamma.cuda(), beta.cuda())
        close(out.view(), ref.view(), 1, True)
    for (B2, H, W, ws, shift) in ((2, 16, 24, 8, 4), (2, 8, 12, 12, 6), (4, 9, 13, 8, 4)):
        C, geo = 96, WinGeom(B2, H, W, ws, shift)
        (xh, xf), ref = h_map(B2, H, W, C, g), f32_out_like(1, 1, geo.rows, C)
        gamma, beta = torch.randn(C, generator=g), torch.randn(C, generator=g)
        em.window_gather_ln(xf, ref, geo, gamma, beta)
        out = Map(torch.zeros(1, 1, geo.rows, C, dtype=torch.float16).cuda(), 0, C)
        cu.window_gather_ln(to_gpu(xh), out, geo, gamma.cuda(), beta.cuda())
        close(out.view(), ref.view(), 1, True)
    for (B, H, W, C) in ((2, 9, 13, 448), (1, 40, 37, 224), (2, 5, 16, 104), (1, 70, 8, 64)):
        (xh, xf), ref = h_map(B, H, W, C, g), f32_out_like(B, H, W, C)
        P = {"d.weight": torch.randn(C, 1, 3, 3, generator=g) * 0.4, "d.bias": torch.randn(C, generator=g)}
        w9c, b = pack.pack_dw(P, "d")
        em.dwconv_gelu(xf, ref, w9c, b)
        out = Map(torch.zeros(B, H, W, (C + 7) // 8 * 8, dtype=torch.float16).cuda(), 0, C)
        cu.dwconv_gelu(to_gpu(xh), out, w9c.cuda(), b.cuda())
        close(out.view(), ref.view(), 9, True)
    # token warp: fp16 features, fp32 flows
    B, H, W, C = 2, 34, 60, 96
    (xh, xf), ref = h_map(B, H, W, C, g), f32_out_like(B, H, W, C)
    head = Map(torch.randn(B, H, W, 8, generator=g) * 7.0, 0, 5)
    em.flow_warp_nhwc(xf, head, 2, ref)
    out = Map(torch.zeros(B, H, W, C, dtype=torch.float16).cuda(), 0, C)
    cu.flow_warp_nhwc(to_gpu(xh), to_gpu(head), 2, out)
    close(out.view(), ref.view(), 4, True)
    # first conv and image packer write fp16
    img = torch.rand(2, 3, 21, 30, generator=g)
    P = {"c.weight": torch.randn(24, 3, 3, 3, generator=g) * 0.3, "c.bias": torch.randn(24, generator=g) * 0.1, "p": torch.rand(24, generator=g) * 0.5}
    w = pack.pack_conv(P, "c", prelu="p")
    ref = f32_out_like(2, 21, 30, 24)
    em.conv3x3_first(img, w, ref)
    out = Map(torch.zeros(2, 21, 30, 24, dtype=torch.float16).cuda(), 0, 24)
    cu.conv3x3_first(img.cuda(), _pg_to_gpu(w), out)
    close(out.view(), ref.view(), 27, True)
    imgs = [torch.rand(2, 3, 21, 30, generator=g) for _ in range(5)]
    ref = f32_out_like(2, 21, 30, 16)
    em.pack5_planar(imgs, Map(ref.t, 0, 15))
    out = Map(torch.zeros(2, 21, 30, 16, dtype=torch.float16).cuda(), 0, 15)
    cu.pack5_planar([t.cuda() for t in imgs], out)
    close(out.t, ref.t, 1, True)
    m = Map(torch.randn(2, 10, 12, 8, generator=g), 0, 5)
    t = cu.to_act(to_gpu(m))
    assert t.half and t.pitch == 8 and torch.equal(t.view().cpu(), m.view().half()) and float(t.t[..., 5:].abs().max()) == 0


# end to end: fp16 has the mantissa of TF32 -> the tf32 tolerances
for _v in ("default", "stress", "ensemble"):
    TOL[("f16", _v)] = TOL[("tf32", _v)]


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[5:-4] for p in CASES])
def test_f16_forward_matches_reference_golden(path):
    z = np.load(path)
    meta = json.loads(str(z["meta"]))
    P = weights.make_weights(meta["kind"], meta["variant"])
    im0, im1 = weights.synthetic_frames(meta["B"], meta["H"], meta["W"], kind=meta["frames"])
    net = _net(meta["kind"], P)
    net.global_motion = meta["global_motion"]
    net.ensemble_global_motion = bool(meta.get("ensemble", False))
    net.precision = "f16"
    out = net(im0.cuda(), im1.cuda())
    tol = TOL[("f16", meta["variant"])]
    errs = {k: float(np.abs(out[k].cpu().numpy() - z[k]).max()) for k in ("I_t", "I_t_0", "I_t_1", "occ_mask1", "opt_flow_0", "opt_flow_1")}
    print(f"[f16 golden] {os.path.basename(path)}: {errs} mean {float(np.abs(out['I_t'].cpu().numpy() - z['I_t']).mean()):.3e}")
    for key in ("I_t", "I_t_0", "I_t_1", "occ_mask1"):
        assert errs[key] <= tol["img"], (key, errs[key])
    for key in ("opt_flow_0", "opt_flow_1"):
        assert errs[key] <= tol["flow"], (key, errs[key])
    assert np.abs(out["I_t"].cpu().numpy() - z["I_t"]).mean() <= tol["mean"]
    if meta["variant"] == "default":
        assert psnr(out["I_t"].cpu(), torch.from_numpy(z["I_t"])) >= 60


def test_f16_full_size_1080p_and_psnr_delta():
    """configs[2] at full size in the fp16 mode: tf32 tolerances, dPSNR versus ground truth <= 0.01 dB; and the uint8 API."""
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    P = weights.make_weights("base", "default")
    im0, gt, im1 = weights.synthetic_triplet(1, 1088, 1920)
    ref = oracle.forward(P, im0, im1, True)
    net = _net("base", P)
    net.precision = "f16"
    out = net(im0.cuda(), im1.cuda())
    e = (out["I_t"].cpu() - ref["I_t"]).abs()
    fe = max((out[k].cpu() - ref[k]).abs().max().item() for k in ("opt_flow_0", "opt_flow_1"))
    p_new, p_ref = psnr(out["I_t"].cpu(), gt), psnr(ref["I_t"], gt)
    print(f"[1080p default] f16: max|I_t| {e.max().item():.3e} mean {e.mean().item():.3e}, max|flow| {fe:.3e} px, PSNR(new, ref) {psnr(out['I_t'].cpu(), ref['I_t']):.1f} dB; "
          f"PSNR vs ground truth: ref {p_ref:.4f} dB, new {p_new:.4f} dB, delta {p_new - p_ref:+.5f} dB")
    tol = TOL[("f16", "default")]
    assert e.max().item() <= tol["img"] and fe <= tol["flow"] and psnr(out["I_t"].cpu(), ref["I_t"]) >= 60
    assert abs(p_new - p_ref) <= 0.01
    del net
    torch.cuda.empty_cache()
