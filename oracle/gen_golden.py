"""Generate tests/golden/* by running the UNMODIFIED reference (build container only).

    python oracle/gen_golden.py            # rewrites tests/golden/{schema_*.json, case_*.npz}

Each case loads a deterministic weight set (oracle/weights.py) into the real reference ``Network`` with
``strict=True``, runs ``forward`` on seeded synthetic frames on CPU fp32, and stores the frames' seed
and the reference outputs.  Tests then compare the oracle (and, on the GPU, the CUDA path) with these.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import refshim  # noqa: E402
import weights  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")

# name, kind, variant, B, H, W, global, frames
CASES = [
    ("lite_default_off_96x128", "lite", "default", 1, 96, 128, False, "noise"),
    ("lite_stress_on_128x192", "lite", "stress", 1, 128, 192, True, "texture"),
    ("base_default_on_128x192", "base", "default", 1, 128, 192, True, "noise"),
    ("base_stress_on_b2_64x96", "base", "stress", 2, 64, 96, True, "texture"),
    ("lite_stress_off_b2_72x104", "lite", "stress", 2, 72, 104, False, "texture"),
    # forward_global_ensemble (network_base.py:617-712): ensemble_global_motion=True, H and W multiples of 64
    ("ensemble_base_default_128x128", "base", "default", 1, 128, 128, True, "noise"),
    ("ensemble_select_lite_b3_128x192", "lite", "ensemble", 3, 128, 192, True, "shift"),   # a different scale wins per sample
]


def run_case(name, kind, variant, b, h, w, glob, frames):
    Net = refshim.load_reference_network(kind)
    P = weights.make_weights(kind, variant)
    ens = name.startswith("ensemble")
    net = Net(global_motion=glob, ensemble_global_motion=ens).eval()
    net.load_state_dict(P, strict=True)
    im0, im1 = weights.synthetic_frames(b, h, w, kind=frames)
    with torch.no_grad():
        out = net(im0, im1)
        if ens:      # also record which scale won, for a meaningful test (the selection must not be degenerate)
            g0, g1 = net.multiscale_global_motion_ensemble(im0, im1)
    rec = {"I_t": out["I_t"], "opt_flow_0": out["opt_flow_0"], "opt_flow_1": out["opt_flow_1"],
           "occ_mask1": out["occ_mask1"], "I_t_0": out["I_t_0"], "I_t_1": out["I_t_1"]}
    for i, t in enumerate(out["im_t_list"]):
        rec[f"im_t_list_{i}"] = t
    rec["coarse_im0_warped"] = out["im0_warped_list"][-1]
    rec["coarse_im1_warped"] = out["im1_warped_list"][-1]
    if ens:
        rec["ensemble_flow_0"], rec["ensemble_flow_1"] = g0, g1
    np.savez_compressed(os.path.join(GOLDEN, f"case_{name}.npz"),
                        meta=json.dumps(dict(kind=kind, variant=variant, B=b, H=h, W=w, global_motion=glob, frames=frames, ensemble=ens)),
                        **{k: v.numpy().astype(np.float32) for k, v in rec.items()})
    f0 = out["opt_flow_0"]
    print(f"{name}: I_t mean {out['I_t'].mean():.4f} std {out['I_t'].std():.4f} clamp% "
          f"{((out['I_t']==0)|(out['I_t']==1)).float().mean()*100:.1f} | flow0 absmax {f0.abs().max():.3f} std {f0.std():.3f} "
          f"| occ [{out['occ_mask1'].min():.3f},{out['occ_mask1'].max():.3f}]")


def run_config0():
    """BASELINE.json configs[0]: Lite, 2x interpolation of asset/example_frame0/1.png, global motion off, CPU fp32,
    through the reference's own InputPadder + forward + rounding (demo_2x.py:54-87 minus the .cuda() calls)."""
    import cv2
    import torch.nn.functional as F
    Net = refshim.load_reference_network("lite")
    net = Net(global_motion=False).eval()
    net.load_state_dict(weights.make_weights("lite", "default"), strict=True)
    a = cv2.imread(os.path.join(refshim.REF_ROOT, "asset", "example_frame0.png"))
    b = cv2.imread(os.path.join(refshim.REF_ROOT, "asset", "example_frame1.png"))
    sys.path.insert(0, refshim.REF_ROOT)
    refshim.install()
    import importlib.util
    spec = importlib.util.spec_from_file_location("_ref_bench_utils", os.path.join(refshim.REF_ROOT, "benchmark", "utils.py"))
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    sys.path.pop(0)
    t0 = (torch.tensor(a[:, :, ::-1].copy().transpose(2, 0, 1)) / 255.).unsqueeze(0)
    t1 = (torch.tensor(b[:, :, ::-1].copy().transpose(2, 0, 1)) / 255.).unsqueeze(0)
    padder = mod.InputPadder(t0.shape, divisor=64)
    p0, p1 = padder.pad(t0, t1)
    with torch.no_grad():
        pred = padder.unpad(net(p0, p1)["I_t"][0])
    pred = np.round(pred.numpy().transpose(1, 2, 0) * 255).astype(np.uint8)[:, :, ::-1].copy()
    np.savez_compressed(os.path.join(GOLDEN, "case_config0_lite_example_frames.npz"),
                        meta=json.dumps(dict(kind="lite", variant="default", B=1, H=a.shape[0], W=a.shape[1], global_motion=False, frames="asset")),
                        pred_bgr=pred)
    print("config0:", pred.shape, "mean", pred.mean())


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    if not sys.argv[1:] or "config0" in sys.argv[1:]:
        run_config0()
    for kind in ("base", "lite"):
        Net = refshim.load_reference_network(kind)
        torch.manual_seed(0)
        sd = Net().state_dict()
        with open(os.path.join(GOLDEN, f"schema_{kind}.json"), "w") as f:
            json.dump({k: list(v.shape) for k, v in sd.items()}, f, indent=0)
    only = sys.argv[1:]
    for c in CASES:
        if not only or c[0] in only:
            run_case(*c)


if __name__ == "__main__":
    main()
