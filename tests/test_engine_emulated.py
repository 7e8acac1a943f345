"""Host-logic check (CPU): the product's packing (pack.py) and orchestration (engine.py), driven through
the CPU emulation of the operator contracts (tests/emul_ops.py), must reproduce the oracle."""
import pytest
import torch

import atmvfi_oracle as oracle
import weights
from atmvfi.arch import ARCHS
from atmvfi.engine import PackedModel, Plan
from emul_ops import EmulOps

CASES = [
    ("lite", "stress", 1, 64, 96, False),
    ("lite", "stress", 1, 128, 192, True),      # 1/16 grid 8x12 -> padded 12x12: pad mask + shift mask
    ("base", "stress", 2, 64, 96, True),
    ("lite", "default", 2, 72, 104, False),     # 1/8 grid 9x13 -> padded 16x16
]


@pytest.mark.parametrize("kind,variant,B,H,W,glob", CASES)
def test_plan_matches_oracle(kind, variant, B, H, W, glob):
    P = weights.make_weights(kind, variant)
    im0, im1 = weights.synthetic_frames(B, H, W, kind="texture")
    ref = oracle.forward(P, im0, im1, glob)
    model = PackedModel(ARCHS[kind], P, 8, 12, with_global=glob)
    plan = Plan(EmulOps(), model, B, H, W, glob)
    out = plan.run(im0, im1)
    tol = 5e-3 if variant == "stress" else 2e-5
    for key in ("I_t", "opt_flow_0", "opt_flow_1", "occ_mask1", "occ_mask2", "I_t_0", "I_t_1"):
        assert out[key].shape == ref[key].shape, key
        err = (out[key] - ref[key]).abs().max().item()
        assert err <= tol, (key, err)
    for name in ("im_t_list", "im0_warped_list", "im1_warped_list"):
        assert len(out[name]) == len(ref[name]) == (5 if glob else 4)
        for a, b in zip(out[name], ref[name]):
            assert a.shape == b.shape
            assert (a - b).abs().max().item() <= tol, name


def test_bad_shape_raises():
    P = weights.make_weights("lite", "default")
    model = PackedModel(ARCHS["lite"], P, 8, 12)
    with pytest.raises(RuntimeError):
        Plan(EmulOps(), model, 1, 72, 104, True)    # not a multiple of 16 with global motion


@pytest.mark.parametrize("kind,variant,B,H,W,frames", [("lite", "ensemble", 3, 128, 192, "shift"), ("base", "default", 1, 128, 128, "noise")])
def test_ensemble_plan_matches_oracle(kind, variant, B, H, W, frames):
    """forward_global_ensemble (network_base.py:564-712): 3-scale global motion + per-sample selection."""
    P = weights.make_weights(kind, variant)
    im0, im1 = weights.synthetic_frames(B, H, W, kind=frames)
    ref = oracle.forward(P, im0, im1, True, ensemble=True)
    model = PackedModel(ARCHS[kind], P, 8, 12, with_global=True)
    plan = Plan(EmulOps(), model, B, H, W, True, ensemble=True)
    out = plan.run(im0, im1)
    if variant == "ensemble":      # a different scale wins for every sample
        losses = torch.stack([l.reshape(-1) for l in plan.ensemble_losses], 1)
        assert losses.argmin(1).tolist() == [0, 1, 2]
    tol = 5e-3 if variant != "default" else 2e-5
    for key in ("I_t", "opt_flow_0", "opt_flow_1", "occ_mask1"):
        assert (out[key] - ref[key]).abs().max().item() <= tol, key
    assert len(out["im_t_list"]) == len(ref["im_t_list"]) == 4
    with pytest.raises(RuntimeError):
        Plan(EmulOps(), model, 1, 96, 128, True, ensemble=True)     # 96 is not a multiple of 64


def test_head_major_qkv_layout_matches_oracle():
    """The fused q|k|v linear writing the head-major layout (ATMVFI_OUT_QKV_HEADS) + attention reading it: same forward."""
    P = weights.make_weights("lite", "stress")
    im0, im1 = weights.synthetic_frames(1, 128, 192, kind="texture")
    ref = oracle.forward(P, im0, im1, True)
    model = PackedModel(ARCHS["lite"], P, 8, 12, with_global=True)
    out = Plan(EmulOps(qkv_head_major=True), model, 1, 128, 192, True).run(im0, im1)
    for key in ("I_t", "opt_flow_0", "opt_flow_1", "occ_mask1"):
        assert (out[key] - ref[key]).abs().max().item() <= 5e-3, key


def test_stream_plan_reuses_encoder_features():
    """Video-stream plan: the encoder features of frame k+1 are computed once (as frame 1 of pair k) and copied to the frame-0 half
    for pair k+1; every pair must equal the ordinary two-frame forward."""
    P = weights.make_weights("lite", "stress")
    model = PackedModel(ARCHS["lite"], P, 8, 12, with_global=True)
    frames = [weights.synthetic_frames(1, 64, 96, seed=s, kind="texture")[0] for s in (1, 2, 3, 4)]
    normal = Plan(EmulOps(), model, 1, 64, 96, True)
    stream = Plan(EmulOps(), model, 1, 64, 96, True, stream=True)
    assert len(stream.records) < len(normal.records) + 4 and stream.encode_records
    stream.im1.copy_(frames[0])
    stream.encode_only()
    for k in range(len(frames) - 1):
        ref = {kk: (v.clone() if torch.is_tensor(v) else v) for kk, v in normal.run(frames[k], frames[k + 1]).items()}
        stream.im0.copy_(stream.im1)
        stream.im1.copy_(frames[k + 1])
        out = stream.run_inplace()
        for key in ("I_t", "opt_flow_0", "opt_flow_1", "occ_mask1"):
            assert (out[key] - ref[key]).abs().max().item() <= 1e-4, (k, key)     # stress gains x batch-1 vs batch-2 conv summation order on CPU
    with pytest.raises(NotImplementedError):
        Plan(EmulOps(), model, 1, 64, 64, True, ensemble=True, stream=True)
