"""SURVEY 8(f) rows 3 and 4 on the B200: the evaluation-harness arithmetic (InputPadder, flip TTA - test_xiph.py:128-140,
test_snufilm.py:118-131) and the recursive 4x / 8x interpolation (davis-vid.py:102-106) against the CPU oracle chained the way the
reference scripts chain the reference model."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

import atmvfi_oracle as oracle
import weights
from test_gpu_forward import _net


def _oracle_mid(P, a, b, glob=True):
    return oracle.forward(P, a, b, glob)["I_t"]


@pytest.mark.parametrize("divisor,shape", [(32, (100, 180)), (64, (90, 150))])
def test_flip_tta_prediction_matches_chained_oracle(divisor, shape):
    from benchmark.harness import predict_middle
    P = weights.make_weights("lite", "default")
    H, W = shape
    im0, im1 = weights.synthetic_frames(1, H, W, kind="texture")
    # the reference harness: replicate-pad to the divisor, forward, forward of the doubly flipped pair, average, un-pad
    eh, ew = (-H) % divisor, (-W) % divisor
    pad = [ew // 2, ew - ew // 2, eh // 2, eh - eh // 2]
    a, b = F.pad(im0, pad, mode="replicate"), F.pad(im1, pad, mode="replicate")
    ref = _oracle_mid(P, a, b)
    ref_tta = (ref + _oracle_mid(P, a.flip(2).flip(3), b.flip(2).flip(3)).flip(2).flip(3)) / 2
    crop = lambda t: t[..., pad[2]: t.shape[-2] - pad[3], pad[0]: t.shape[-1] - pad[1]]
    net = _net("lite", P)
    net.precision = "fp32"
    got = predict_middle(net, im0.cuda(), im1.cuda(), divisor=divisor, TTA=False).cpu()
    got_tta = predict_middle(net, im0.cuda(), im1.cuda(), divisor=divisor, TTA=True).cpu()
    assert got.shape == (1, 3, H, W)
    assert (got - crop(ref)).abs().max().item() <= 1e-4
    assert (got_tta - crop(ref_tta)).abs().max().item() <= 1e-4


@pytest.mark.parametrize("levels", [2, 3])
def test_recursive_interpolation_matches_chained_oracle(levels):
    """4x (the reference's INTERPOLATE4X) and 8x: every level feeds on the fp32 output of the level above."""
    P = weights.make_weights("lite", "default")
    im0, im1 = weights.synthetic_frames(1, 128, 192, kind="texture")

    def rec(a, b, d):
        m = _oracle_mid(P, a, b)
        return [m] if d == 1 else rec(a, m, d - 1) + [m] + rec(m, b, d - 1)

    want = rec(im0, im1, levels)
    net = _net("lite", P)
    net.precision = "fp32"
    got = net.interpolate_recursive(im0.cuda(), im1.cuda(), levels=levels)
    assert len(got) == 2 ** levels - 1
    for i, (g, w) in enumerate(zip(got, want)):
        assert (g.cpu() - w).abs().max().item() <= 2e-4, i
    # central frame with the script's flip TTA
    tta = net.interpolate_recursive(im0.cuda(), im1.cuda(), levels=levels, TTA=True)
    c = len(want) // 2
    want_c = (want[c] + _oracle_mid(P, im0.flip(2).flip(3), im1.flip(2).flip(3)).flip(2).flip(3)) / 2
    assert (tta[c].cpu() - want_c).abs().max().item() <= 2e-4
    assert all(torch.equal(tta[i], got[i]) for i in range(len(got)) if i != c)


def test_recursive_u8_and_davis_loop():
    """uint8 front end: the 4x frames equal inference_2frame arithmetic applied to the chained fp32 frames (NOT to re-quantised ones)."""
    from demo_2x import inference_2frame, inference_multiframe
    from benchmark.davis_vid import interpolate_sequence
    P = weights.make_weights("lite", "default")
    net = _net("lite", P)
    rng = np.random.default_rng(3)
    f = [rng.integers(0, 256, (70, 100, 3), dtype=np.uint8) for _ in range(5)]
    mids = inference_multiframe(f[0], f[2], net, levels=2)
    assert len(mids) == 3 and all(m.shape == f[0].shape and m.dtype == np.uint8 for m in mids)
    assert np.array_equal(mids[1], inference_2frame(f[0], f[2], net))          # the central frame is the plain 2x result
    # quarter frames come from the fp32 middle frame: close to, but not the same bytes as, the result of chaining through uint8
    # (random-noise frames: a re-quantised input moves single pixels by a few LSB)
    q = inference_2frame(f[0], mids[1], net)
    d = np.abs(q.astype(int) - mids[0].astype(int))
    assert d.mean() <= 1.5 and d.max() <= 24
    big = [np.ascontiguousarray(np.pad(x, ((5, 5), (6, 6), (0, 0)), mode="edge")) for x in f]
    out = list(interpolate_sequence(net, big, time_interval=2, H=64, W=96))
    assert len(out) == 2 * 4 + 1 and out[0].shape == big[0].shape and out[1].shape == (64, 96, 3)
