"""Pins the CPU oracle (oracle/atmvfi_oracle.py) to outputs of the UNMODIFIED reference.

tests/golden/case_*.npz were produced by oracle/gen_golden.py, which runs the real reference
(/root/reference, network_base.py / network_lite.py) on CPU fp32 with the deterministic weight sets of
oracle/weights.py.  The reference itself has no tests or golden vectors.
"""
import glob
import json
import os

import numpy as np
import pytest
import torch

import atmvfi_oracle as oracle
import weights

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
CASES = [p for p in sorted(glob.glob(os.path.join(GOLDEN, "case_*.npz"))) if "config0" not in p]


def test_golden_present():
    assert len(CASES) >= 4


@pytest.mark.parametrize("kind", ["base", "lite"])
def test_schema_matches_reference(kind):
    with open(os.path.join(GOLDEN, f"schema_{kind}.json")) as f:
        ref = {k: tuple(v) for k, v in json.load(f).items()}
    mine = dict(weights.schema(kind))
    assert set(mine) == set(ref)
    assert len(ref) == 236
    for k in ref:
        assert mine[k] == ref[k], k


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[5:-4] for p in CASES])
def test_oracle_matches_reference_output(path):
    z = np.load(path)
    meta = json.loads(str(z["meta"]))
    P = weights.make_weights(meta["kind"], meta["variant"])
    im0, im1 = weights.synthetic_frames(meta["B"], meta["H"], meta["W"], kind=meta["frames"])
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    ens = bool(meta.get("ensemble", False))
    out = oracle.forward(P, im0, im1, meta["global_motion"], ensemble=ens)
    # same ATen kernels, same op order -> agreement to fp32 round-off (thread-count noise floor 8.7e-6,
    # multiplied by the stress gains)
    tol = 2e-5 if meta["variant"] == "default" else 5e-3
    for key in ("I_t", "opt_flow_0", "opt_flow_1", "occ_mask1", "I_t_0", "I_t_1"):
        err = np.abs(out[key].numpy() - z[key]).max()
        assert err <= tol, (key, err)
    n = len([k for k in z.files if k.startswith("im_t_list_")])
    assert len(out["im_t_list"]) == n == (5 if meta["global_motion"] and not ens else 4)
    if ens:      # the 3-scale selection itself (network_base.py:564-615)
        g0, g1, losses = oracle.multiscale_global_motion_ensemble(P, im0, im1, oracle.window_sizes(P)[1])
        assert np.abs(g0.numpy() - z["ensemble_flow_0"]).max() <= tol and np.abs(g1.numpy() - z["ensemble_flow_1"]).max() <= tol
    for i in range(n):
        err = np.abs(out["im_t_list"][i].numpy() - z[f"im_t_list_{i}"]).max()
        assert err <= tol, (i, err)
    assert np.abs(out["im0_warped_list"][-1].numpy() - z["coarse_im0_warped"]).max() <= tol


def test_oracle_inference_2frame_config0():
    """BASELINE config 0 through the oracle's restatement of demo_2x.inference_2frame, byte for byte."""
    import cv2
    a = cv2.imread(os.path.join(GOLDEN, "example_frame0.png"))
    b = cv2.imread(os.path.join(GOLDEN, "example_frame1.png"))
    z = np.load(os.path.join(GOLDEN, "case_config0_lite_example_frames.npz"))
    P = weights.make_weights("lite", "default")
    pred = oracle.inference_2frame(P, a, b, global_motion=False, is_bgr=True)
    assert pred.shape == (600, 414, 3)
    assert np.array_equal(pred, z["pred_bgr"])
