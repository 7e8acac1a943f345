"""Per-operator parity of the tcgen05 (kind::tf32) implicit-GEMM kernel - the kernel that carries 3/4 of the forward - against
the TF32-operand emulation of its contract (tests/emul_ops.py, EmulOps(tf32=True)): activations truncated and weights rounded to
TF32 exactly as the tensor core / pack_tc do, products accumulated in fp64.  With the operand rounding emulated the comparison
is tight (2e-5 on O(1) outputs, fp32 accumulation order only), so a misplaced tap, a wrong channel tail, a swapped source or a
bad row remap cannot hide behind the 1e-3 of TF32 itself.  Shapes cover every kernel variant: halo boxes, paired tiles,
2-CTA clusters, 8 / 16 epilogue warps, stride 2 / 4, dilation, 2-4 concatenated sources, channel tails (101, 197, 389, 773),
several N tiles, ConvTranspose pixel-shuffle, dual output, residual, window reverse, head-major q|k|v."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from atmvfi import _lib, pack
from atmvfi.ops import CudaOps, Map, WinGeom, PackedGemm
from emul_ops import EmulOps, round_tf32
from gpu_util import max_err, to_gpu

TOL = 2e-5


def tol_k(K):
    """Bound for a layer with K products per output (outputs are O(1)).  The tensor core accumulates in fp32 with truncation
    when it aligns addends (measured on B200: the error against exact accumulation grows linearly, ~1.3e-8 per product:
    4.5e-5 at K = 3501, 9.3e-5 at K = 6957), so the bound follows K; a misplaced tap or channel costs >= 1e-2."""
    return max(TOL, 2e-8 * K)


@pytest.fixture(scope="module")
def ops():
    assert torch.cuda.is_available()
    cu = CudaOps(torch.device("cuda:0"), _lib.TF32)
    cu.round_outputs = False                 # compare the accumulators; the rounding of stored maps is checked on its own below
    return cu, EmulOps(tf32=True)


def gen(seed=0):
    return torch.Generator().manual_seed(seed)


def tf32_map(B, H, W, C, g, pitch=None, scale=1.0):
    """Random NHWC map whose values are already TF32 (what every producer stores in this mode)."""
    pitch = pitch or (C + 3) // 4 * 4
    return Map(round_tf32(torch.randn(B, H, W, pitch, generator=g) * scale), 0, C)


def _pg_to_gpu(w: PackedGemm) -> PackedGemm:
    c = lambda t: None if t is None else t.cuda()
    return PackedGemm(w.name, w.ksize, w.split, w.Cout, w.shuffle, c(w.w32), c(w.bias), c(w.prelu))


def _conv_weights(ci, co, k, g):
    return {"c.weight": torch.randn(co, ci, k, k, generator=g) / (ci * k * k) ** 0.5, "c.bias": torch.randn(co, generator=g) * 0.1,
            "p": torch.rand(co, generator=g) * 0.5}


TC_CONV_CASES = [
    # B, H, W, splits, Cout, k, stride, dil
    (1, 32, 40, [24], 48, 3, 2, 1),                    # encoder stride 2
    (2, 32, 48, [48], 48, 3, 4, 1),                    # fusion stride 4
    (2, 32, 48, [48], 48, 3, 4, 2),                    # fusion stride 4, dilation 2
    (1, 9, 13, [16, 20, 20], 36, 3, 1, 1),             # halo boxes on an odd, tiny grid (tiles hang over every border)
    (1, 16, 16, [96, 48, 48, 192], 384, 1, 1, 1),      # 4 sources, two N tiles
    (1, 20, 28, [101, 15], 64, 3, 1, 1),               # channel tails, paired tiles (N <= 128)
    (1, 8, 12, [64], 5, 1, 1, 1),                      # motion head: Cout = 5
    (1, 130, 70, [16], 3, 3, 1, 1),                    # refine head: Cout = 3, tall grid
    (1, 40, 56, [389], 197, 3, 1, 1),                  # decoder widths, N = 208, 2-CTA cluster
    (2, 24, 40, [773], 389, 3, 1, 1),                  # two N tiles, K = 6975
    (1, 64, 96, [24], 24, 3, 1, 1),                    # thin full-resolution layer
    (1, 33, 47, [197, 48], 101, 3, 1, 1),              # skip concat with tails on odd sizes
    (3, 17, 23, [32], 64, 1, 1, 1),                    # short-K 1x1: 16 epilogue warps
    (1, 68, 120, [384], 384, 1, 1, 1),                 # transformer linear at the 1080p 1/16 grid
]


@pytest.mark.parametrize("B,H,W,split,Co,k,stride,dil", TC_CONV_CASES)
@pytest.mark.parametrize("act", [True, False])
def test_tc_conv(ops, B, H, W, split, Co, k, stride, dil, act):
    cu, em = ops
    g = gen(11)
    w = pack.pack_conv(_conv_weights(sum(split), Co, k, g), "c", split=split, prelu="p")
    srcs = [tf32_map(B, H, W, c, g, pitch=(c + 3) // 4 * 4 + 4 * (i % 2)) for i, c in enumerate(split)]
    pad = dil * (k - 1) // 2
    Ho = (H + 2 * pad - dil * (k - 1) - 1) // stride + 1
    Wo = (W + 2 * pad - dil * (k - 1) - 1) // stride + 1
    out_c = Map(torch.zeros(B, Ho, Wo, (Co + 3) // 4 * 4 + 4), 4 if Co % 4 == 0 else 0, Co)
    out_g = to_gpu(out_c)
    em.gemm_conv(srcs, w, out_c, stride=stride, dil=dil, act=act)
    rec = cu.recording = []
    cu.gemm_conv(to_gpu(srcs), _pg_to_gpu(w), out_g, stride=stride, dil=dil, act=act)
    cu.recording = None
    assert rec[0][3][0].precision == _lib.TF32, "the layer fell back to the CUDA-core kernel: this test must exercise tcgen05"
    cu.replay(rec)
    assert max_err(out_g, out_c) < tol_k(k * k * sum(split))
    if out_g.c0:
        assert out_g.t.cpu()[..., : out_g.c0].abs().max() == 0        # channels outside the written slice stay untouched


@pytest.mark.parametrize("split,Co,H,W", [([37], 21, 7, 9), ([16, 16], 16, 7, 9), ([384, 384, 5], 37, 7, 9), ([389], 197, 34, 60), ([64, 64], 32, 40, 33)])
def test_tc_transposed(ops, split, Co, H, W):
    cu, em = ops
    g = gen(12)
    ci = sum(split)
    P = {"d.0.weight": torch.randn(ci, Co, 2, 2, generator=g) / ci ** 0.5, "d.0.bias": torch.randn(Co, generator=g) * 0.1,
         "d.1.weight": torch.rand(Co, generator=g) * 0.5}
    w = pack.pack_deconvp(P, "d", split=split)
    srcs = [tf32_map(2, H, W, c, g) for c in split]
    out = tf32_map(2, 2 * H, 2 * W, Co, g)
    og = to_gpu(out)
    em.gemm_conv(srcs, w, out)
    cu.gemm_conv(to_gpu(srcs), _pg_to_gpu(w), og)
    assert max_err(og, out) < TOL


def test_tc_dual_output_residual_and_rounding(ops):
    cu, em = ops
    g = gen(13)
    # decoder level: raw output + PReLU'd copy (engine: gemm_conv(..., act=False, out2=act, prelu2=nxt))
    for (ci, co, H, W) in ((40, 29, 12, 10), (197, 101, 40, 56)):
        w = pack.pack_conv(_conv_weights(ci, co, 3, g), "c")
        src = tf32_map(2, H, W, ci, g)
        slopes = torch.rand(co, generator=g)
        o1, o2 = tf32_map(2, H, W, co, g), tf32_map(2, H, W, co, g)
        g1, g2 = to_gpu(o1), to_gpu(o2)
        em.gemm_conv([src], w, o1, act=False, out2=o2, prelu2=slopes)
        cu.gemm_conv([to_gpu(src)], _pg_to_gpu(w), g1, act=False, out2=g2, prelu2=slopes.cuda())
        assert max_err(g1, o1) < tol_k(9 * ci) and max_err(g2, o2) < tol_k(9 * ci)
    # linear + residual (attention proj / Mlp fc2): the residual-prefetch epilogue
    for (ci, co, rows) in ((40, 52, 333), (384, 384, 4100), (1536, 384, 700)):
        Pl = {"l.weight": torch.randn(co, ci, generator=g) / ci ** 0.5, "l.bias": torch.randn(co, generator=g)}
        wl = pack.pack_linear(Pl, ["l"])
        x, res, out = tf32_map(1, 1, rows, ci, g), tf32_map(1, 1, rows, co, g), tf32_map(1, 1, rows, co, g)
        og = to_gpu(out)
        em.gemm_conv([x], wl, out, act=False, residual=res)
        cu.gemm_conv([to_gpu(x)], _pg_to_gpu(wl), og, act=False, residual=to_gpu(res))
        assert max_err(og, out) < tol_k(ci), (ci, co, rows)
    # stored maps rounded to TF32 (the production setting): representable in 10 mantissa bits and within one TF32 ulp
    cu.round_outputs = True
    try:
        w = pack.pack_conv(_conv_weights(48, 64, 3, g), "c", prelu="p")
        src, o = tf32_map(1, 30, 44, 48, g), tf32_map(1, 30, 44, 64, g)
        og = to_gpu(o)
        em.gemm_conv([src], w, o)
        cu.gemm_conv([to_gpu(src)], _pg_to_gpu(w), og)
        got = og.view().cpu()
        assert (got.contiguous().view(torch.int32) & 0x1FFF).abs().max().item() == 0
        assert ((got - o.view()).abs() <= o.view().abs() * 2 ** -10 + 1e-6).all()
    finally:
        cu.round_outputs = False


WIN_CASES = [(2, 16, 24, 8, 0), (2, 16, 24, 8, 4), (4, 9, 13, 8, 4), (2, 8, 12, 12, 6), (2, 8, 12, 12, 0), (2, 68, 120, 12, 6)]


@pytest.mark.parametrize("B2,H,W,ws,shift", WIN_CASES)
def test_tc_window_reverse(ops, B2, H, W, ws, shift):
    """Attention projection: linear + residual on the normed window rows + window reverse / un-roll / de-pad (attention.py:320-331)."""
    cu, em = ops
    g = gen(14)
    C = 96
    geo = WinGeom(B2, H, W, ws, shift)
    Pl = {"l.weight": torch.randn(C, C, generator=g) * 0.1, "l.bias": torch.randn(C, generator=g)}
    wl = pack.pack_linear(Pl, ["l"])
    x, res = tf32_map(1, 1, geo.rows, C, g), tf32_map(1, 1, geo.rows, C, g)
    out = tf32_map(B2, H, W, C, g)
    og = to_gpu(out)
    em.gemm_conv([x], wl, out, act=False, residual=res, win=geo)
    cu.gemm_conv([to_gpu(x)], _pg_to_gpu(wl), og, act=False, residual=to_gpu(res), win=geo)
    assert max_err(og, out) < TOL


@pytest.mark.parametrize("hd,rows", [(48, 1000), (84, 577), (28, 130), (44, 4096)])
def test_tc_qkv_head_major(ops, hd, rows):
    cu, em = ops
    g = gen(15)
    heads, C = 8, 8 * hd
    Pl = {"q.weight": torch.randn(C, C, generator=g) / C ** 0.5, "kv.weight": torch.randn(2 * C, C, generator=g) / C ** 0.5}
    wl = pack.pack_linear(Pl, ["q", "kv"], bias=False)
    x = tf32_map(1, 1, rows, C, g)
    out = Map(torch.zeros(1, 1, rows, 3 * C), 0, 3 * C)
    og = to_gpu(out)
    em.gemm_conv([x], wl, out, act=False, qkv_heads=heads)
    cu.gemm_conv([to_gpu(x)], _pg_to_gpu(wl), og, act=False, qkv_heads=heads)
    assert max_err(og, out) < TOL


# ---------------------------------------------------------------------------------------------------------------------
# 3xTF32 ("fp32x3"): fp32-tolerance products on the tensor cores.  Inputs are arbitrary fp32 values (NOT pre-rounded) and the
# reference is the plain fp32 contract emulation: the same bound as the CUDA-core FFMA kernel's tests (tests/test_gpu_ops.py).
# ---------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def ops3():
    return CudaOps(torch.device("cuda:0"), _lib.TF32X3), EmulOps()


def f32_map(B, H, W, C, g, pitch=None):
    pitch = pitch or (C + 3) // 4 * 4
    return Map(torch.randn(B, H, W, pitch, generator=g), 0, C)


@pytest.mark.parametrize("B,H,W,split,Co,k,stride,dil", TC_CONV_CASES)
def test_x3_conv(ops3, B, H, W, split, Co, k, stride, dil):
    cu, em = ops3
    g = gen(21)
    w = pack.pack_conv(_conv_weights(sum(split), Co, k, g), "c", split=split, prelu="p")
    srcs = [f32_map(B, H, W, c, g, pitch=(c + 3) // 4 * 4 + 4 * (i % 2)) for i, c in enumerate(split)]
    pad = dil * (k - 1) // 2
    Ho = (H + 2 * pad - dil * (k - 1) - 1) // stride + 1
    Wo = (W + 2 * pad - dil * (k - 1) - 1) // stride + 1
    out_c = Map(torch.zeros(B, Ho, Wo, (Co + 3) // 4 * 4 + 4), 4 if Co % 4 == 0 else 0, Co)
    out_g = to_gpu(out_c)
    em.gemm_conv(srcs, w, out_c, stride=stride, dil=dil)
    rec = cu.recording = []
    cu.gemm_conv(to_gpu(srcs), _pg_to_gpu(w), out_g, stride=stride, dil=dil)
    cu.recording = None
    assert rec[0][3][0].precision == _lib.TF32X3
    cu.replay(rec)
    assert max_err(out_g, out_c) < tol_k(k * k * sum(split))


def test_x3_transposed_residual_window(ops3):
    cu, em = ops3
    g = gen(22)
    for split, Co, H, W in (([37], 21, 7, 9), ([384, 384, 5], 37, 7, 9), ([389], 197, 34, 60)):
        ci = sum(split)
        P = {"d.0.weight": torch.randn(ci, Co, 2, 2, generator=g) / ci ** 0.5, "d.0.bias": torch.randn(Co, generator=g) * 0.1,
             "d.1.weight": torch.rand(Co, generator=g) * 0.5}
        w = pack.pack_deconvp(P, "d", split=split)
        srcs = [f32_map(2, H, W, c, g) for c in split]
        out = f32_map(2, 2 * H, 2 * W, Co, g)
        og = to_gpu(out)
        em.gemm_conv(srcs, w, out)
        cu.gemm_conv(to_gpu(srcs), _pg_to_gpu(w), og)
        assert max_err(og, out) < tol_k(sum(split)), (split, Co)
    for (ci, co, rows) in ((40, 52, 333), (384, 384, 4100), (1536, 384, 700)):
        Pl = {"l.weight": torch.randn(co, ci, generator=g) / ci ** 0.5, "l.bias": torch.randn(co, generator=g)}
        wl = pack.pack_linear(Pl, ["l"])
        x, res, out = f32_map(1, 1, rows, ci, g), f32_map(1, 1, rows, co, g), f32_map(1, 1, rows, co, g)
        og = to_gpu(out)
        em.gemm_conv([x], wl, out, act=False, residual=res)
        cu.gemm_conv([to_gpu(x)], _pg_to_gpu(wl), og, act=False, residual=to_gpu(res))
        assert max_err(og, out) < tol_k(ci), (ci, co, rows)
    geo = WinGeom(2, 68, 120, 12, 6)
    Pl = {"l.weight": torch.randn(96, 96, generator=g) * 0.1, "l.bias": torch.randn(96, generator=g)}
    wl = pack.pack_linear(Pl, ["l"])
    x, res, out = f32_map(1, 1, geo.rows, 96, g), f32_map(1, 1, geo.rows, 96, g), f32_map(2, 68, 120, 96, g)
    og = to_gpu(out)
    em.gemm_conv([x], wl, out, act=False, residual=res, win=geo)
    cu.gemm_conv([to_gpu(x)], _pg_to_gpu(wl), og, act=False, residual=to_gpu(res), win=geo)
    assert max_err(og, out) < TOL
    # dual output
    w = pack.pack_conv(_conv_weights(197, 101, 3, g), "c")
    src, slopes = f32_map(2, 40, 56, 197, g), torch.rand(101, generator=g)
    o1, o2 = f32_map(2, 40, 56, 101, g), f32_map(2, 40, 56, 101, g)
    g1, g2 = to_gpu(o1), to_gpu(o2)
    em.gemm_conv([src], w, o1, act=False, out2=o2, prelu2=slopes)
    cu.gemm_conv([to_gpu(src)], _pg_to_gpu(w), g1, act=False, out2=g2, prelu2=slopes.cuda())
    assert max_err(g1, o1) < tol_k(9 * 197) and max_err(g2, o2) < tol_k(9 * 197)
