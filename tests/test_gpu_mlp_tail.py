"""Fused Mlp tail (atmvfi_mlp_tail: DWConv3x3 + GELU produced in shared memory as the A operand of fc2, + bias + residual) against
(a) the two stand-alone launches it replaces (atmvfi_dwconv3x3_gelu + atmvfi_gemm_conv with a residual) - same operand bits, same
K order, so the results agree to the accumulation order of the tensor core - and (b) the CPU contract emulation of those two
operators (reference: attention.py:74-85, 118-123, 333).  Shapes: the Base / Lite token grids of the local (1/8) and global (1/16)
branches, grids whose tiles hang over the border, odd tile counts (phantom tile of the CTA pair), one and two N tiles."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from atmvfi import _lib, pack
from atmvfi.ops import CudaOps, Map
from emul_ops import EmulOps, round_tf32
from gpu_util import max_err, to_gpu

CASES = [
    # B2, H, W, C, hidden
    (2, 16, 32, 384, 1536),       # Base local widths, whole tiles
    (2, 17, 30, 384, 1536),       # tiles hang over both borders
    (1, 8, 16, 384, 1536),        # a single tile: the second CTA of the pair runs a phantom tile
    (2, 9, 20, 672, 2688),        # Base global widths: two N tiles of 352
    (2, 12, 24, 224, 448),        # Lite local: one MMA of N = 224
    (1, 10, 18, 352, 704),        # Lite global: N = 352 = 256 + 96
    (2, 34, 60, 384, 1536),       # 1080p / 4 in each direction: several tiles per CTA, both accumulator phases
]


def _weights(C, hid, g):
    P = {"fc2.weight": torch.randn(C, hid, generator=g) / hid ** 0.5, "fc2.bias": torch.randn(C, generator=g) * 0.1,
         "dw.weight": torch.randn(hid, 1, 3, 3, generator=g) / 3.0, "dw.bias": torch.randn(hid, generator=g) * 0.1}
    return pack.pack_linear(P, ["fc2"]), pack.pack_dw(P, "dw")


def _pg_to_gpu(w):
    from atmvfi.ops import PackedGemm
    c = lambda t: None if t is None else t.cuda()
    return PackedGemm(w.name, w.ksize, w.split, w.Cout, w.shuffle, c(w.w32), c(w.bias), c(w.prelu))


@pytest.mark.parametrize("B,H,W,C,hid", CASES)
@pytest.mark.parametrize("precision", ["tf32", "f16"])
def test_mlp_tail_matches_unfused_and_emulation(B, H, W, C, hid, precision):
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(5)
    f16 = precision == "f16"
    cu = CudaOps(dev, _lib.F16 if f16 else _lib.TF32)
    fc2, (dw_w, dw_b) = _weights(C, hid, g)
    h = torch.randn(B, H, W, hid, generator=g)
    x = torch.randn(B, H, W, C, generator=g)
    if f16:
        h, x = h.half().float(), x.half().float()
    else:
        h, x = round_tf32(h), round_tf32(x)
    dt = torch.float16 if f16 else torch.float32
    hg, xg = Map(h.to(dev, dt), 0, hid), Map(x.to(dev, dt), 0, C)
    fc2g, dw_wg, dw_bg = _pg_to_gpu(fc2), dw_w.cuda(), dw_b.cuda()
    # (a) the two stand-alone launches
    h2 = cu.new_map(B, H, W, hid)
    cu.dwconv_gelu(hg, h2, dw_wg, dw_bg)
    ref = cu.new_map(B, H, W, C)
    cu.gemm_conv([h2.rows()], fc2g, ref.rows(), act=False, residual=xg.rows())
    # fused
    out = cu.new_map(B, H, W, C)
    out.t.fill_(float("nan"))
    assert cu.mlp_tail_ok(hg, fc2g, xg, out)
    cu.mlp_tail(hg, dw_wg, dw_bg, fc2g, xg, out)
    torch.cuda.synchronize()
    assert torch.isfinite(out.t.float()).all()
    e_unfused = max_err(out, ref)
    # (b) CPU emulation of the two operators
    em = EmulOps(tf32=not f16)
    eh2 = Map(torch.zeros(B, H, W, hid), 0, hid)
    em.dwconv_gelu(Map(h, 0, hid), eh2, dw_w, dw_b)
    if f16:
        eh2 = Map(eh2.t.half().float(), 0, hid)
        w = fc2.w32.clone()
        fc2e = type(fc2)(fc2.name, fc2.ksize, fc2.split, fc2.Cout, fc2.shuffle, w.half().float(), fc2.bias, fc2.prelu)
        em = EmulOps()
    else:
        eh2 = Map(round_tf32(eh2.t), 0, hid)
        fc2e = fc2
    eo = Map(torch.zeros(B, H, W, C), 0, C)
    em.gemm_conv([eh2.rows()], fc2e, eo.rows(), act=False, residual=Map(x, 0, C).rows())
    e_emul = max_err(out, eo)
    print(f"mlp_tail {precision} {B}x{H}x{W} C={C} hid={hid}: vs unfused {e_unfused:.3e}, vs emulation {e_emul:.3e}")
    # outputs are O(1): same operand bits and K order as the unfused launches (fp16 stores: one ulp of 2^-10 at |v| <= 4)
    assert e_unfused <= (4e-3 if f16 else 2e-5), e_unfused
    assert e_emul <= (8e-3 if f16 else 4e-3), e_emul       # + the TF32 rounding of the stored output (2^-11 relative)


@pytest.mark.parametrize("precision", ["tf32", "f16"])
def test_mlp_tail_row_window(precision):
    """Row window (row-slab mode): the rows of the window equal the full launch bit for bit, rows outside it are not touched, and
    hidden rows beyond the window's one-row halo are never needed (they are poisoned with NaN here)."""
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(9)
    f16 = precision == "f16"
    B, H, W, C, hid = 2, 29, 40, 384, 1536
    cu = CudaOps(dev, _lib.F16 if f16 else _lib.TF32)
    fc2, (dw_w, dw_b) = _weights(C, hid, g)
    dt = torch.float16 if f16 else torch.float32
    h = (torch.randn(B, H, W, hid, generator=g).half().float() if f16 else round_tf32(torch.randn(B, H, W, hid, generator=g))).to(dev, dt)
    x = (torch.randn(B, H, W, C, generator=g).half().float() if f16 else round_tf32(torch.randn(B, H, W, C, generator=g))).to(dev, dt)
    fc2g, dw_wg, dw_bg = _pg_to_gpu(fc2), dw_w.cuda(), dw_b.cuda()
    full = cu.new_map(B, H, W, C)
    cu.mlp_tail(Map(h, 0, hid), dw_wg, dw_bg, fc2g, Map(x, 0, C), full)
    for y0, y1 in ((0, 9), (9, 21), (21, 29), (5, 6)):
        hp = h.clone()
        hp[:, : max(0, y0 - 1)] = float("nan")
        hp[:, y1 + 1 :] = float("nan")
        part = cu.new_map(B, H, W, C)
        part.t.fill_(7.0)
        cu.mlp_tail(Map(hp, 0, hid), dw_wg, dw_bg, fc2g, Map(x, 0, C), part, rows=(y0, y1))
        torch.cuda.synchronize()
        assert torch.equal(part.t[:, y0:y1], full.t[:, y0:y1]), (y0, y1)
        assert (part.t[:, :y0] == 7.0).all() and (part.t[:, y1:] == 7.0).all(), (y0, y1)


@pytest.mark.parametrize("kind,precision", [("base", "tf32"), ("lite", "tf32"), ("base", "f16")])
def test_forward_is_bit_identical_with_and_without_the_fused_tail(kind, precision, monkeypatch):
    """Whole forward (six transformer blocks, local + global branch): the plan with atmvfi_mlp_tail and the plan with the two
    stand-alone launches (ATMVFI_MLP_TAIL=0) produce the same bits in every output."""
    import weights
    from test_gpu_forward import _net
    P = weights.make_weights(kind, "stress")
    im0, im1 = weights.synthetic_frames(1, 192, 256, kind="texture")
    outs = []
    for flag in ("1", "0"):
        monkeypatch.setenv("ATMVFI_MLP_TAIL", flag)
        net = _net(kind, P)
        net.precision = precision
        o = net(im0.cuda(), im1.cuda())
        torch.cuda.synchronize()
        names = {r[0] for r in next(iter(net._runtime._plans.values())).records}
        assert ("atmvfi_mlp_tail" in names) == (flag == "1") and ("atmvfi_dwconv3x3_gelu" in names) == (flag == "0")
        flat = {}
        for k, v in o.items():
            for j, t in enumerate(v if isinstance(v, (list, tuple)) else [v]):
                if isinstance(t, torch.Tensor):
                    flat[f"{k}[{j}]"] = t.clone()
        outs.append(flat)
        del net
    assert len(outs[0]) >= 10 and outs[0].keys() == outs[1].keys()
    for k in outs[0]:
        assert torch.equal(outs[0][k], outs[1][k]), k
