"""Per-operator parity on the B200: every C-ABI kernel against the CPU emulation of its contract
(tests/emul_ops.py, stock torch ops) on seeded inputs, including the edge cases the reference exercises:
window padding + shift, flows that leave the image, strided / dilated / concatenated-input convolutions."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from atmvfi import _lib, pack
from atmvfi.ops import CudaOps, Map, WinGeom, PackedGemm
from emul_ops import EmulOps
from gpu_util import max_err, rand_map, to_gpu


@pytest.fixture(scope="module")
def ops():
    assert torch.cuda.is_available()
    return CudaOps(torch.device("cuda:0"), _lib.FP32), EmulOps()


def gen(seed=0):
    return torch.Generator().manual_seed(seed)


def _pg_to_gpu(w: PackedGemm) -> PackedGemm:
    c = lambda t: None if t is None else t.cuda()
    return PackedGemm(w.name, w.ksize, w.split, w.Cout, w.shuffle, c(w.w32), c(w.bias), c(w.prelu))


def _conv_weights(ci, co, k, g, prelu=True):
    P = {"c.weight": torch.randn(co, ci, k, k, generator=g) / (ci * k * k) ** 0.5, "c.bias": torch.randn(co, generator=g) * 0.1,
         "p": torch.rand(co, generator=g) * 0.5}
    return P


CONV_CASES = [
    # B, H, W, splits, Cout, k, stride, dil
    (2, 17, 23, [3], 24, 3, 1, 1),
    (1, 32, 40, [24], 48, 3, 2, 1),
    (2, 32, 48, [48], 48, 3, 4, 1),
    (2, 32, 48, [48], 48, 3, 4, 2),
    (1, 9, 13, [8, 20, 20], 36, 3, 1, 1),
    (1, 16, 16, [96, 48, 48, 192], 384, 1, 1, 1),
    (1, 20, 28, [101, 15], 64, 3, 1, 1),
    (1, 8, 12, [64], 5, 1, 1, 1),
    (1, 130, 70, [16], 3, 3, 1, 1),
]


@pytest.mark.parametrize("B,H,W,split,Co,k,stride,dil", CONV_CASES)
def test_gemm_conv_plain(ops, B, H, W, split, Co, k, stride, dil):
    cu, em = ops
    g = gen(1)
    P = _conv_weights(sum(split), Co, k, g)
    w = pack.pack_conv(P, "c", split=split, prelu="p")
    srcs = [rand_map(B, H, W, c, pitch=(c + 3) // 4 * 4 + 4 * (i % 2), gen=g) for i, c in enumerate(split)]
    pad = dil * (k - 1) // 2
    Ho = (H + 2 * pad - dil * (k - 1) - 1) // stride + 1
    Wo = (W + 2 * pad - dil * (k - 1) - 1) // stride + 1
    out_c = Map(torch.zeros(B, Ho, Wo, (Co + 3) // 4 * 4 + 4), 4 if Co % 4 == 0 else 0, Co)
    out_g = to_gpu(out_c)
    em.gemm_conv(srcs, w, out_c, stride=stride, dil=dil)
    cu.gemm_conv(to_gpu(srcs), _pg_to_gpu(w), out_g, stride=stride, dil=dil)
    assert max_err(out_g, out_c) < 2e-5
    # channels outside the written slice stay untouched
    if out_g.c0:
        assert out_g.t.cpu()[..., : out_g.c0].abs().max() == 0


def test_gemm_conv_dual_output_and_residual(ops):
    cu, em = ops
    g = gen(2)
    P = _conv_weights(40, 29, 3, g)
    w = pack.pack_conv(P, "c")
    src = rand_map(2, 12, 10, 40, gen=g)
    slopes = torch.rand(29, generator=g)
    o1, o2 = rand_map(2, 12, 10, 29, gen=g), rand_map(2, 12, 10, 29, gen=g)
    g1, g2 = to_gpu(o1), to_gpu(o2)
    em.gemm_conv([src], w, o1, act=False, out2=o2, prelu2=slopes)
    cu.gemm_conv([to_gpu(src)], _pg_to_gpu(w), g1, act=False, out2=g2, prelu2=slopes.cuda())
    assert max_err(g1, o1) < 2e-5 and max_err(g2, o2) < 2e-5
    # linear + residual
    Pl = {"l.weight": torch.randn(52, 40, generator=g) * 0.1, "l.bias": torch.randn(52, generator=g)}
    wl = pack.pack_linear(Pl, ["l"])
    x, res, out = rand_map(1, 1, 333, 40, gen=g), rand_map(1, 1, 333, 52, gen=g), rand_map(1, 1, 333, 52, gen=g)
    og = to_gpu(out)
    em.gemm_conv([x], wl, out, act=False, residual=res)
    cu.gemm_conv([to_gpu(x)], _pg_to_gpu(wl), og, act=False, residual=to_gpu(res))
    assert max_err(og, out) < 2e-5


@pytest.mark.parametrize("split,Co", [([37], 21), ([12, 12], 16), ([384, 384, 5], 37)])
def test_gemm_conv_transposed(ops, split, Co):
    cu, em = ops
    g = gen(3)
    ci = sum(split)
    P = {"d.0.weight": torch.randn(ci, Co, 2, 2, generator=g) / ci ** 0.5, "d.0.bias": torch.randn(Co, generator=g) * 0.1,
         "d.1.weight": torch.rand(Co, generator=g) * 0.5}
    w = pack.pack_deconvp(P, "d", split=split)
    srcs = [rand_map(2, 7, 9, c, gen=g) for c in split]
    out = rand_map(2, 14, 18, Co, gen=g)
    og = to_gpu(out)
    em.gemm_conv(srcs, w, out)
    cu.gemm_conv(to_gpu(srcs), _pg_to_gpu(w), og)
    assert max_err(og, out) < 2e-5


WIN_CASES = [(2, 16, 24, 8, 0), (2, 16, 24, 8, 4), (4, 9, 13, 8, 4), (2, 8, 12, 12, 6), (2, 8, 12, 12, 0), (2, 10, 7, 4, 2)]


@pytest.mark.parametrize("B2,H,W,ws,shift", WIN_CASES)
def test_window_gather_ln_and_reverse(ops, B2, H, W, ws, shift):
    cu, em = ops
    g = gen(4)
    C = 32
    geo = WinGeom(B2, H, W, ws, shift)
    tok = rand_map(B2, H, W, C, gen=g)
    gamma, beta = torch.randn(C, generator=g), torch.randn(C, generator=g)
    win = rand_map(1, 1, geo.rows, C, gen=g)
    wg = to_gpu(win)
    em.window_gather_ln(tok, win, geo, gamma, beta)
    cu.window_gather_ln(to_gpu(tok), wg, geo, gamma.cuda(), beta.cuda())
    assert max_err(wg, win) < 1e-5
    # projection + residual + window reverse
    Pl = {"l.weight": torch.randn(C, C, generator=g) * 0.2, "l.bias": torch.randn(C, generator=g)}
    wl = pack.pack_linear(Pl, ["l"])
    x = rand_map(1, 1, geo.rows, C, gen=g)
    out = rand_map(B2, H, W, C, gen=g)
    og = to_gpu(out)
    em.gemm_conv([x], wl, out, act=False, residual=win, win=geo)
    cu.gemm_conv([to_gpu(x)], _pg_to_gpu(wl), og, act=False, residual=to_gpu(win), win=geo)
    assert max_err(og, out) < 2e-5


@pytest.mark.parametrize("B2,H,W,ws,shift", WIN_CASES)
@pytest.mark.parametrize("hd", [28, 48])
def test_window_attention_motion(ops, B2, H, W, ws, shift, hd):
    cu, em = ops
    g = gen(5)
    heads, C = 8, 8 * hd
    geo = WinGeom(B2, H, W, ws, shift)
    qkv = rand_map(1, 1, geo.rows, 3 * C, gen=g, scale=1.5)
    N = ws * ws
    idx = torch.arange(N)
    px, py = (idx % ws).float(), (idx // ws).float()
    rc = torch.stack([px[None] - px[:, None], py[None] - py[:, None]], 0).contiguous()
    mix = (torch.randn(4, 8, generator=g), torch.randn(4, generator=g), torch.randn(4, generator=g), torch.randn(1, generator=g))
    out, mo = rand_map(1, 1, geo.rows, C, gen=g), rand_map(B2 // 2, H, W, 8, gen=g)
    og, mg = to_gpu(out), to_gpu(mo)
    scratch = torch.empty(geo.rows * heads * 2, device="cuda")
    em.window_attention(qkv, out, geo, heads, True, rc, mix, mo, 4)
    cu.window_attention(to_gpu(qkv), og, geo, heads, True, rc.cuda(), to_gpu(mix), mg, 4, scratch)
    assert max_err(og, out) < 2e-5
    assert max_err(mg, mo) < 1e-4
    # self-attention variant without motion
    em.window_attention(qkv, out, geo, heads, False)
    cu.window_attention(to_gpu(qkv), og, geo, heads, False)
    assert max_err(og, out) < 2e-5


def test_layernorm_dwconv(ops):
    cu, em = ops
    g = gen(6)
    for C in (224, 384, 672):
        x, o = rand_map(1, 3, 50, C, gen=g, scale=3.0), rand_map(1, 3, 50, C, gen=g)
        gamma, beta = torch.randn(C, generator=g), torch.randn(C, generator=g)
        og = to_gpu(o)
        em.layernorm(x, o, gamma, beta)
        cu.layernorm(to_gpu(x), og, gamma.cuda(), beta.cuda())
        assert max_err(og, o) < 2e-5
    # channel counts that are / are not multiples of the TMA kernel's 64-channel block, widths off the 16-column tile,
    # heights beyond one 34-row strip
    for (B, H, W, C) in ((2, 9, 13, 448), (1, 40, 37, 224), (2, 5, 16, 100), (1, 70, 8, 64)):
        x, o = rand_map(B, H, W, C, gen=g), rand_map(B, H, W, C, gen=g)
        P = {"d.weight": torch.randn(C, 1, 3, 3, generator=g) * 0.4, "d.bias": torch.randn(C, generator=g)}
        w9c, b = pack.pack_dw(P, "d")
        og = to_gpu(o)
        em.dwconv_gelu(x, o, w9c, b)
        cu.dwconv_gelu(to_gpu(x), og, w9c.cuda(), b.cuda())
        assert max_err(og, o) < 2e-5, (B, H, W, C)
        # row window: only rows [3, H-1) are produced, everything else keeps its previous contents
        og2 = to_gpu(rand_map(B, H, W, C, gen=g))
        before = og2.t.clone()
        cu.dwconv_gelu(to_gpu(x), og2, w9c.cuda(), b.cuda(), rows=(3, H - 1))
        assert torch.equal(og2.t[:, 3 : H - 1], og.t[:, 3 : H - 1]) and torch.equal(og2.t[:, :3], before[:, :3]) and torch.equal(og2.t[:, H - 1 :], before[:, H - 1 :])


@pytest.mark.parametrize("H,W,mag", [(17, 29, 3.0), (68, 120, 40.0), (136, 240, 300.0)])
def test_warps(ops, H, W, mag):
    """Backward warp incl. flows that leave the frame (zero padding) - flow_warp.py:50-60."""
    cu, em = ops
    g = gen(7)
    B = 2
    im0, im1 = torch.rand(B, 3, H, W, generator=g), torch.rand(B, 3, H, W, generator=g)
    head = rand_map(B, H, W, 5, pitch=12, gen=g, scale=mag)
    head = Map(head.t, 4, 5)
    outs_c = [torch.zeros(B, 3, H, W) for _ in range(3)] + [torch.zeros(B, 2, H, W), torch.zeros(B, 2, H, W), torch.zeros(B, 1, H, W), torch.zeros(B, 1, H, W)]
    outs_g = to_gpu(outs_c)
    em.warp_blend(im0, im1, head, *outs_c)
    cu.warp_blend(im0.cuda(), im1.cuda(), to_gpu(head), *outs_g)
    # tolerance: the fp32 normalise/un-normalise round trip moves coordinates by ~1e-4 px at this size;
    # CPU and GPU agree on that arithmetic, what is left is summation order
    for a, b in zip(outs_g, outs_c):
        assert max_err(a, b) < 5e-5
    flow = torch.randn(B, 2, H, W, generator=g) * mag
    img = torch.rand(B, 7, H, W, generator=g)
    oc = torch.zeros(B, 7, H, W); og = oc.cuda()
    em.flow_warp_nchw(img, flow, oc)
    cu.flow_warp_nchw(img.cuda(), flow.cuda(), og)
    assert max_err(og, oc) < 5e-5
    src, out = rand_map(B, H, W, 48, gen=g), rand_map(B, H, W, 48, gen=g)
    og = to_gpu(out)
    em.flow_warp_nhwc(src, head, 2, out)
    cu.flow_warp_nhwc(to_gpu(src), to_gpu(head), 2, og)
    assert max_err(og, out) < 2e-4 * max(1.0, src.t.abs().max().item() / 4)


def test_resize_pack_finish(ops):
    cu, em = ops
    g = gen(8)
    x = torch.rand(2, 3, 64, 96, generator=g)
    for (h, w, s) in ((32, 48, 1.0), (128, 192, 2.0)):
        oc = torch.zeros(2, 3, h, w); og = oc.cuda()
        em.resize(x, oc, s); cu.resize(x.cuda(), og, s)
        assert max_err(og, oc) < 1e-5
    m = Map(torch.zeros(2, 64, 96, 16), 0, 15); mg = to_gpu(m)
    em.nchw_to_nhwc(x, m.chan(6, 3)); cu.nchw_to_nhwc(x.cuda(), mg.chan(6, 3))
    assert max_err(mg, m) == 0
    res = rand_map(2, 64, 96, 3, gen=g, scale=2.0)
    it = torch.rand(2, 3, 64, 96, generator=g)
    a, b = torch.zeros_like(it), torch.zeros_like(it); ag, bg = a.cuda(), b.cuda()
    em.residual_finish(res, it, a, b); cu.residual_finish(to_gpu(res), it.cuda(), ag, bg)
    assert max_err(ag, a) < 1e-6 and max_err(bg, b) < 1e-6
    assert bg.min() >= 0 and bg.max() <= 1


def test_u8_roundtrip(ops):
    cu, _ = ops
    g = gen(9)
    H, W, Hp, Wp, top, left = 37, 50, 64, 64, 13, 7
    img = torch.randint(0, 256, (H, W, 3), generator=g, dtype=torch.uint8)
    planar = torch.zeros(3, Hp, Wp, device="cuda")
    cu.u8_to_planar(img.cuda(), planar, H, W, Hp, Wp, top, left, True)
    ref = torch.nn.functional.pad((img.flip(-1).permute(2, 0, 1).float() / 255.)[None], (left, Wp - W - left, top, Hp - H - top), mode="replicate")[0]
    assert max_err(planar, ref) == 0
    back = torch.zeros(H, W, 3, dtype=torch.uint8, device="cuda")
    cu.planar_to_u8(planar, back, H, W, Hp, Wp, top, left, True)
    assert torch.equal(back.cpu(), img)


def test_errors_are_loud(ops):
    cu, _ = ops
    x = rand_map(1, 1, 8, 30)        # C=30 not a multiple of 4
    with pytest.raises(_lib.AtmvfiError):
        cu.layernorm(to_gpu(x), to_gpu(x), torch.ones(30).cuda(), torch.zeros(30).cuda())
    with pytest.raises(_lib.AtmvfiError):
        CudaOps(torch.device("cpu"))


@pytest.mark.parametrize("Co", [16, 24])
def test_conv3x3_first_and_pack5(ops, Co):
    cu, em = ops
    g = gen(11)
    P = _conv_weights(3, Co, 3, g)
    w = pack.pack_conv(P, "c", prelu="p")
    img = torch.rand(2, 3, 37, 53, generator=g)
    out = rand_map(2, 37, 53, Co, gen=g)
    og = to_gpu(out)
    em.conv3x3_first(img, w, out)
    cu.conv3x3_first(img.cuda(), _pg_to_gpu(w), og)
    assert max_err(og, out) < 1e-5
    imgs = [torch.rand(2, 3, 37, 53, generator=g) for _ in range(5)]
    m = Map(torch.full((2, 37, 53, 16), 7.0), 0, 15)
    mg = to_gpu(m)
    em.pack5_planar(imgs, m)
    cu.pack5_planar([t.cuda() for t in imgs], mg)
    assert torch.equal(mg.t.cpu(), m.t)


@pytest.mark.parametrize("B2,H,W,ws,shift", WIN_CASES + [(2, 14, 21, 7, 3)])
@pytest.mark.parametrize("hd", [28, 44, 48, 84])
@pytest.mark.parametrize("closed", [False, True])
def test_window_attention_tensor_cores(B2, H, W, ws, shift, hd, closed):
    """tcgen05 attention (TF32 operands) against the fp32 contract emulation."""
    em = EmulOps()
    cu = CudaOps(torch.device("cuda:0"), _lib.TF32)
    g = gen(12)
    heads, C = 8, 8 * hd
    geo = WinGeom(B2, H, W, ws, shift)
    qkv = rand_map(1, 1, geo.rows, 3 * C, gen=g, scale=1.0)
    N = ws * ws
    idx = torch.arange(N)
    px, py = (idx % ws).float(), (idx // ws).float()
    rc = torch.stack([px[None] - px[:, None], py[None] - py[:, None]], 0).contiguous()
    mix = (torch.randn(4, 8, generator=g), torch.randn(4, generator=g), torch.randn(4, generator=g), torch.randn(1, generator=g))
    out, mo = rand_map(1, 1, geo.rows, C, gen=g), rand_map(B2 // 2, H, W, 8, gen=g)
    og, mg = to_gpu(out), to_gpu(mo)
    scratch = torch.empty(geo.rows * heads * 2, device="cuda")
    em.window_attention(qkv, out, geo, heads, True, rc, mix, mo, 4)
    cu.window_attention(to_gpu(qkv), og, geo, heads, True, rc.cuda(), to_gpu(mix), mg, 4, scratch, rc_closed_form=closed)
    # logits have std ~ sqrt(hd)*... / sqrt(hd) = 1: TF32 operand rounding moves them by ~1e-3
    assert max_err(og, out) < 6e-3
    assert max_err(mg, mo) < 4e-2
    em.window_attention(qkv, out, geo, heads, False)
    cu.window_attention(to_gpu(qkv), og, geo, heads, False)
    assert max_err(og, out) < 6e-3


def test_ensemble_reduction_and_select():
    """atmvfi_l1_mean / atmvfi_select_min3 / atmvfi_nhwc_to_nchw against their contracts (network_base.py:548-611)."""
    import torch
    from atmvfi import _lib
    from atmvfi.ops import CudaOps, Map
    ops = CudaOps(torch.device("cuda:0"), _lib.FP32)
    g = torch.Generator().manual_seed(3)
    a, b = torch.rand(3, 3, 70, 90, generator=g).cuda(), torch.rand(3, 3, 70, 90, generator=g).cuda()
    out = torch.empty(3, 1, 1, 1, device="cuda")
    scratch = torch.empty(3, 1, 1, 2048, device="cuda")
    ops.l1_mean(a, b, out, scratch)
    ref = (a.double() - b.double()).abs().mean(dim=[1, 2, 3])
    assert (out.reshape(-1).double() - ref).abs().max().item() <= 1e-7
    l = [torch.tensor([0.1, 0.3, 0.2], device="cuda"), torch.tensor([0.1, 0.2, 0.2], device="cuda"), torch.tensor([0.5, 0.2, 0.1], device="cuda")]
    c = [torch.full((3, 2, 4, 5), float(k), device="cuda") for k in range(3)]
    sel = torch.empty(3, 2, 4, 5, device="cuda")
    ops.select3(l, c, sel)
    assert sel[:, 0, 0, 0].tolist() == [0.0, 1.0, 2.0]          # ties go to the first minimum (if / elif / else chain)
    m = Map(torch.rand(2, 6, 7, 8, generator=g).cuda(), 0, 5)
    pl = torch.empty(2, 2, 6, 7, device="cuda")
    ops.nhwc_to_nchw(m.chan(2, 2), pl)
    assert torch.equal(pl, m.t[..., 2:4].permute(0, 3, 1, 2))


@pytest.mark.parametrize("precision", [_lib.FP32, _lib.TF32], ids=["fp32", "tf32"])
@pytest.mark.parametrize("R,C,heads", [(2 * 64 * 6, 384, 8), (2 * 144 * 2, 224, 8), (1000, 96, 8)])
def test_gemm_qkv_head_major(precision, R, C, heads):
    """Fused q|k|v linear written in the head-major layout (ATMVFI_OUT_QKV_HEADS) against the contract emulation."""
    em = EmulOps()
    cu = CudaOps(torch.device("cuda:0"), precision)
    g = gen(21)
    P = {"l.weight": torch.randn(3 * C, C, generator=g) * 0.05, "l.bias": torch.randn(3 * C, generator=g)}
    w = pack.pack_linear(P, ["l"])
    x = rand_map(1, 1, R, C, gen=g)
    out = rand_map(1, 1, R, 3 * C, gen=g)
    og = to_gpu(out)
    em.gemm_conv([x], w, out, act=False, qkv_heads=heads)
    cu.gemm_conv([to_gpu(x)], _pg_to_gpu(w), og, act=False, qkv_heads=heads)
    assert max_err(og, out) < (2e-5 if precision == _lib.FP32 else 5e-3)


@pytest.mark.parametrize("B2,H,W,ws,shift,hd", [(2, 16, 24, 8, 0, 48), (2, 17, 23, 8, 4, 28), (2, 20, 30, 12, 6, 84), (4, 24, 12, 12, 0, 44), (2, 16, 16, 8, 4, 48)])
@pytest.mark.parametrize("precision", [_lib.FP32, _lib.TF32], ids=["simt", "tcgen05"])
def test_window_attention_head_major(B2, H, W, ws, shift, hd, precision):
    """Attention kernels reading the head-major q / k / v^T layout (TMA-fed on the tensor-core path)."""
    from emul_ops import _to_heads
    em = EmulOps()
    cu = CudaOps(torch.device("cuda:0"), precision)
    g = gen(22)
    heads, C = 8, 8 * hd
    geo = WinGeom(B2, H, W, ws, shift)
    qkv = rand_map(1, 1, geo.rows, 3 * C, gen=g, scale=1.0)
    qkv_h = Map(_to_heads(qkv.t.reshape(geo.rows, 3 * C), heads).reshape(1, 1, geo.rows, 3 * C).contiguous())
    N = ws * ws
    idx = torch.arange(N)
    px, py = (idx % ws).float(), (idx // ws).float()
    rc = torch.stack([px[None] - px[:, None], py[None] - py[:, None]], 0).contiguous()
    mix = (torch.randn(4, 8, generator=g), torch.randn(4, generator=g), torch.randn(4, generator=g), torch.randn(1, generator=g))
    out, mo = rand_map(1, 1, geo.rows, C, gen=g), rand_map(B2 // 2, H, W, 8, gen=g)
    og, mg = to_gpu(out), to_gpu(mo)
    scratch = torch.empty(geo.rows * heads * 2, device="cuda")
    em.window_attention(qkv, out, geo, heads, True, rc, mix, mo, 4)
    cu.window_attention(to_gpu(qkv_h), og, geo, heads, True, rc.cuda(), to_gpu(mix), mg, 4, scratch, rc_closed_form=True, head_major=True)
    tol = 2e-4 if precision == _lib.FP32 else 6e-3
    assert max_err(og, out) < tol
    assert max_err(mg, mo) < (1e-3 if precision == _lib.FP32 else 5e-2)


@pytest.mark.parametrize("H,W,mag,smooth", [(34, 60, 3.0, True), (68, 120, 25.0, True), (136, 240, 60.0, False), (40, 56, 400.0, False)])
def test_pyramid_warp_equals_resize_plus_warp(ops, H, W, mag, smooth):
    """The fused level of the global-motion pyramid (x2 flow up-sampling + both backward warps, source tiles staged in shared memory)
    must give the bits of the unfused launches: smooth flows (staged path), divergent flows (global fallback), flows that leave
    the frame, NaN flows."""
    cu, _ = ops
    g = gen(17)
    B = 2
    im0, im1 = torch.rand(B, 3, H, W, generator=g).cuda(), torch.rand(B, 3, H, W, generator=g).cuda()
    for up in (False, True):
        fh, fw = (H // 2, W // 2) if up else (H, W)
        if smooth:      # a coarse field blown up: what the 1/16-grid global flows look like
            lo = torch.randn(B, 2, 5, 7, generator=g) * mag
            mk = lambda: torch.nn.functional.interpolate(lo + torch.randn(B, 2, 5, 7, generator=g), size=(fh, fw), mode="bilinear", align_corners=True).contiguous().cuda()
        else:
            mk = lambda: (torch.randn(B, 2, fh, fw, generator=g) * mag).cuda()
        f0, f1 = mk(), mk()
        f1[0, :, 3, 4] = float("nan")
        # unfused reference launches
        if up:
            u0, u1 = torch.empty(B, 2, H, W, device="cuda"), torch.empty(B, 2, H, W, device="cuda")
            cu.resize(f0, u0, 2.0); cu.resize(f1, u1, 2.0)
        else:
            u0, u1 = f0, f1
        r0, r1 = torch.empty_like(im0), torch.empty_like(im1)
        cu.flow_warp_nchw(im0, u0, r0); cu.flow_warp_nchw(im1, u1, r1)
        o0, o1 = torch.full_like(im0, -7.0), torch.full_like(im1, -7.0)
        g0, g1 = (torch.empty(B, 2, H, W, device="cuda"), torch.empty(B, 2, H, W, device="cuda")) if up else (None, None)
        cu.pyramid_warp(im0, im1, f0, f1, up, o0, o1, g0, g1)
        assert torch.equal(o0, r0) and torch.equal(o1, r1), (up, (o0 - r0).abs().max().item(), (o1 - r1).abs().max().item())
        if up:
            assert torch.equal(g0, u0) and torch.equal(torch.nan_to_num(g1), torch.nan_to_num(u1))
        # row window
        w0, w1 = torch.full_like(im0, -7.0), torch.full_like(im1, -7.0)
        cu.pyramid_warp(im0, im1, f0, f1, up, w0, w1, None, None, rows=(5, H - 3))
        assert torch.equal(w0[:, :, 5 : H - 3], r0[:, :, 5 : H - 3]) and float(w0[:, :, :5].max()) == -7.0 and float(w1[:, :, H - 3 :].max()) == -7.0
