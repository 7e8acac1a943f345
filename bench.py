#!/usr/bin/env python
"""ATM-VFI forward benchmark (driver contract: python bench.py --gpus N --steps K --warmup W [--impl reference]).

Metric (BASELINE.json): interpolated frames per second (= frame pairs per second).  One "step" = one forward of
the hot path over one batch of synthetic frame pairs.  Default workload: Base network, 1080p (1920x1080 padded
to 1088x1920 as demo_2x.inference_2frame does), global motion on, one pair per step.

  value  : whole-job pairs/s with the frames already resident in HBM (CUDA events, CUDA-graph replay of the plan).
  e2e    : the same through the reference-facing API demo_2x.inference_2frame: numpy uint8 HOST frames in,
           numpy uint8 frame out; H2D/D2H copies, colour conversion, padding and rounding inside the timed region.
  roofline / cpu_baseline: see DESIGN.md section "Measurement".

N > 1 (torchrun, one rank per GPU): frame pairs are independent, so every rank interpolates its own pairs with no
data-path collective (weak scaling); NCCL is used only for the barrier and the max-over-ranks of the elapsed time.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "atm-vfi_b200"), os.path.join(ROOT, "atm-vfi_b200", "network")):
    if p not in sys.path:
        sys.path.insert(0, p)

WORKLOADS = {
    # name: (kind, B, H, W, global_motion, description)
    "base_1080p": ("base", 1, 1080, 1920, True, "network_base Network, 1920x1080 synthetic frame pairs (padded to 1088x1920), global_motion on, 1 pair/step"),
    "lite_1080p": ("lite", 1, 1080, 1920, True, "network_lite Network, 1920x1080 synthetic frame pairs, global_motion on, 1 pair/step"),
    "base_vimeo_b32": ("base", 32, 256, 448, True, "network_base Network, Vimeo90K-shape 448x256 synthetic pairs, batch 32, local+global motion"),
    "base_4k": ("base", 1, 2160, 4096, True, "network_base Network, 4096x2160 synthetic frame pairs (padded to 2176x4096), global_motion on, 1 pair/step"),
    "lite_example": ("lite", 1, 600, 414, False, "network_lite Network, example-frame shape 414x600, global_motion off"),
    # BASELINE.json configs[4]: every rank interpolates a contiguous chunk of the clip; e2e goes through demo_2x.interpolate_video
    "stream_1080p": ("base", 1, 1080, 1920, True, "demo_2x video stream: 24->48 fps 1080p synthetic clip (moving texture), contiguous chunks of frame pairs per GPU"),
}
STREAM_WORKLOADS = {"stream_1080p"}


def synthetic_clip(n_frames, H, W, seed):
    """Moving texture: a smooth random field translated by a constant sub-pixel velocity (so flows are non-trivial), uint8 BGR."""
    import numpy as np
    rng = np.random.default_rng(seed)
    big = rng.random((H // 8 + 40, W // 8 + 40, 3)).astype(np.float32)
    import cv2
    big = cv2.resize(big, ((W // 8 + 40) * 8, (H // 8 + 40) * 8), interpolation=cv2.INTER_CUBIC)
    frames = []
    for k in range(n_frames):
        dy, dx = 16 + 3 * k, 16 + 5 * k
        frames.append(np.ascontiguousarray((np.clip(big[dy:dy + H, dx:dx + W], 0, 1) * 255).astype(np.uint8)))
    return frames




def pad64(h, w):
    return h + (-h) % 64, w + (-w) % 64


class ClockSampler:
    def __init__(self, device_index):
        self.samples, self.reasons, self.proc = [], set(), None
        self.idx = device_index

    def start(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            return
        threading.Thread(target=self._read, daemon=True).start()

    def _read(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.proc.stdout:
            f = [x.strip() for x in line.split(",")]
            try:
                self.samples.append((int(f[0]), int(f[1])))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    self.reasons.add(n)

    def reset(self):
        self.samples, self.reasons = [], set()

    def stop(self):
        if self.proc:
            self.proc.kill()
        s = sorted(x[0] for x in self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.samples[0][1] if self.samples else None,
                "reasons": sorted(self.reasons), "samples": len(s)}


def load_peaks():
    """(HBM GB/s, dense bf16 TFLOP/s burst, sustained, basis) - driver-measured, else the profiling recipe's fallback."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            pk = json.load(f)
        return pk["hbm_gbs"], pk["bf16_tflops"], pk["bf16_tflops_sustained"], "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, 1650.0, 1400.0, "fallback (B200_PROFILING.md)"


def measure_matmul_peak(torch, dev, dtype, device_index, seconds=2.0):
    """Dense library GEMM peak of this GPU in this job: torch.matmul 8192^3 (cuBLAS), best of 10 (burst) and back to back for
    `seconds` (sustained), with the SM clock sampled under load.  dtype "tf32": fp32 operands with allow_tf32; "bf16"/"f16"."""
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        n = 8192
        td = {"tf32": torch.float32, "bf16": torch.bfloat16, "f16": torch.float16}[dtype]
        a = torch.randn(n, n, device=dev, dtype=td)
        b = torch.randn(n, n, device=dev, dtype=td)
        c = torch.empty(n, n, device=dev, dtype=td)
        fl = 2.0 * n ** 3
        for _ in range(3):
            torch.matmul(a, b, out=c)
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); torch.matmul(a, b, out=c); e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        sampler = ClockSampler(device_index)
        sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps, t0 = 0, time.time()
        e0.record()
        while time.time() - t0 < seconds:
            for _ in range(10):
                torch.matmul(a, b, out=c)
            reps += 10
            torch.cuda.synchronize()
        e1.record()
        torch.cuda.synchronize()
        clk = sampler.stop()
        sustained = fl * reps / (e0.elapsed_time(e1) * 1e-3) / 1e12
        return {"dtype": dtype, "burst_tflops": round(fl / (best * 1e-3) / 1e12, 1), "sustained_tflops": round(sustained, 1),
                "sm_mhz_sustained": clk["sm_mhz"], "how": f"torch.matmul {n}^3, best of 10 / back to back {seconds:.0f} s, in this job"}
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


def build_net(kind, device, precision):
    import torch
    from network_base import Network as NB
    from network_lite import Network as NL
    torch.manual_seed(0)
    net = (NB if kind == "base" else NL)()          # random-init weights of the named architecture
    net = net.to(device).eval()
    net.precision = precision
    return net


def synthetic_u8(B, H, W, seed):
    import numpy as np
    rng = np.random.default_rng(seed)
    return [(rng.integers(0, 256, (H, W, 3), dtype=np.uint8), rng.integers(0, 256, (H, W, 3), dtype=np.uint8)) for _ in range(B)]


TC_PRECISIONS = {1: "tf32 tcgen05", 2: "3xtf32 tcgen05", 3: "f16 tcgen05"}      # atmvfi_gemm_conv_desc.precision codes on the tensor cores
# MMA rate of each tensor-core datapath relative to dense bf16 (kind::tf32 issues at half the kind::f16 rate; 3xTF32 spends 3 MMAs per product)
TC_RATE_VS_BF16 = {1: 0.5, 2: 0.5 / 3.0, 3: 1.0}


def _rows(y0, y1, H):
    return (y1 - y0) if y1 else H


def algorithmic_bytes(name, args, keep):
    """Bytes one launch of an HBM-bound kernel must move (inputs read once + outputs written once, fp32 unless the map is
    16-bit), from the C-ABI arguments (include/atmvfi.h); None for kernels not in an HBM family.  Figures per unit: DESIGN.md section 4."""
    es = getattr(keep[0], "esize", 4) if keep and hasattr(keep[0], "esize") else 4
    if name == "atmvfi_dwconv3x3_gelu":
        _, _, B, H, W, C, _, _, _, y0, y1 = args[:11]
        return "dwconv3x3+gelu", 2 * B * _rows(y0, y1, H) * W * C * es
    if name == "atmvfi_mlp_tail":
        # fused DWConv + GELU + fc2 + residual: hidden map read once, residual read, output written (the activated hidden map that the
        # two stand-alone launches wrote and re-read - another 2 x hidden bytes - never exists)
        B, H, W, Ch, C, prec, y0, y1 = args[2], args[3], args[4], args[5], args[14], args[15], args[16], args[17]
        e = 2 if prec == 3 else 4
        return "mlp_tail (dwconv+gelu+fc2+res)", B * _rows(y0, y1, H) * W * (Ch + 2 * C) * e
    if name in ("atmvfi_flow_warp_nhwc", "atmvfi_flow_warp_nhwc_p2p"):
        B, C, H, W, y0, y1 = args[7:13]
        return "flow_warp", B * _rows(y0, y1, H) * W * (2 * C * es + 8)
    if name == "atmvfi_flow_warp_nchw":
        b, c, h, w, y0, y1 = args[3:9]
        return "flow_warp", b * _rows(y0, y1, h) * w * (2 * c * 4 + 8)
    if name == "atmvfi_pyramid_warp":
        up, b, h, w, y0, y1 = args[4], args[9], args[10], args[11], args[12], args[13]
        px = b * _rows(y0, y1, h) * w
        flows = (2 * 2 * 4 / 4 if up else 2 * 2 * 4) + sum(8 for a in args[7:9] if a is not None)
        return "flow_warp", int(px * (2 * (3 + 3) * 4 + flows))
    if name in ("atmvfi_warp_blend", "atmvfi_warp_blend_p2p"):
        b, h, w, y0, y1 = args[12:17]
        extra = sum(8 if a is not None else 0 for a in args[8:10]) + sum(4 if a is not None else 0 for a in args[10:12])
        return "flow_warp", b * _rows(y0, y1, h) * w * (24 + 20 + 36 + extra)      # 80 B/pixel for the fused pair warp + blend
    if name in ("atmvfi_window_attention", "atmvfi_window_attention_tc"):
        C = args[4]
        g = keep[2]
        nrows = g.B2 * g.Hp * g.Wp
        y0, y1 = args[-3], args[-2]
        if y1:
            nrows = g.B2 * (y1 - y0) * g.ws * g.Wp
        return "window_attention", nrows * 4 * C * 4            # q | k | v read, o written
    if name == "atmvfi_layernorm":
        return "layernorm", args[4] * 2 * args[5] * es
    if name == "atmvfi_window_gather_ln":
        g = keep[2]
        return "layernorm", g.B2 * (g.H * g.W + g.Hp * g.Wp) * args[4] * es
    if name == "atmvfi_resize_bilinear_ac":
        _, _, BC, Hin, Win, Hout, Wout, _, y0, y1 = args[:10]
        return "resize", BC * (Hin * Win + _rows(y0, y1, Hout) * Wout) * 4
    if name == "atmvfi_conv3x3_first":
        b, h, w, co, y0, y1 = args[7:13]
        return "conv3x3_first", b * _rows(y0, y1, h) * w * (3 * 4 + co * es)
    return None


def measure_kernels(plan, torch):
    """Per-launch device time of every record of the plan (CUDA events on the launching stream), grouped by entry
    point; returns (table, tc_summary, hbm_families) where tc_summary describes the dominant tcgen05 GEMM kernel and
    hbm_families the HBM-bound kernel families (algorithmic bytes / measured time)."""
    ops = plan.ops
    recs = plan.records
    ops.set_rounding()
    st = torch.cuda.current_stream().cuda_stream
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(recs) + 1)]
    for rep in range(2):                      # first pass warms caches / clocks
        ev[0].record()
        for i, (name, fn, args, _) in enumerate(recs):
            fn(*args, st)
            ev[i + 1].record()
        torch.cuda.synchronize()
    per, fam = {}, {}
    tc_flops = tc_ms = 0.0
    tc_n = 0
    tc_prec = None
    for i, (name, fn, args, keep) in enumerate(recs):
        ms = ev[i].elapsed_time(ev[i + 1])
        key = name
        if name == "atmvfi_gemm_conv":
            d = keep[0]
            key = f"atmvfi_gemm_conv[{TC_PRECISIONS[d.precision]}]" if d.precision in TC_PRECISIONS else "atmvfi_gemm_conv[fp32 ffma]"
            if d.precision in TC_PRECISIONS:
                cin = sum(d.src[s].C for s in range(d.nsrc))
                n = d.Cout * (4 if d.out_mode == 1 else 1)
                hrows = (d.row_end - d.row_begin) if d.row_end else d.Hout          # row window of a slab plan
                tc_flops += 2.0 * d.B * hrows * d.Wout * cin * d.ksize * d.ksize * n
                tc_ms += ms
                tc_n += 1
                tc_prec = d.precision
        else:
            ab = algorithmic_bytes(name, args, keep)
            if ab is not None:
                f = fam.setdefault(ab[0], [0, 0.0, 0.0])
                f[0] += 1; f[1] += ab[1]; f[2] += ms
        a = per.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += ms
    return per, (tc_flops, tc_ms, tc_n, tc_prec), fam


def roofline_records(per, tc, fam, workload, peaks, total_ms):
    """The `roofline` object of the dominant kernel + `roofline_hbm` for the HBM-bound kernel families."""
    hbm_gbs, bf16_burst, bf16_sust, basis, tf32_meas = peaks
    tc_flops, tc_ms, tc_n, tc_prec = tc
    roof = None
    if tc_n:
        traffic = None
        for tp in ("r02_tc_dram_traffic.json", "r01_tc_dram_traffic.json"):
            tp = os.path.join(ROOT, "profiles", tp)
            if os.path.exists(tp):
                try:
                    traffic = json.load(open(tp)).get(workload)
                except Exception:
                    traffic = None
                break
        achieved = tc_flops / tc_n / (tc_ms / tc_n * 1e-3) / 1e12
        rate = TC_RATE_VS_BF16[tc_prec]
        # per-launch times come from a short replay at boost clocks -> the BURST peak is the denominator.  For kind::tf32 the
        # peak is the larger of the library TF32 GEMM measured in this job and half the driver-measured bf16 burst figure.
        peak = bf16_burst * rate
        note = f"{basis}: dense bf16 burst {bf16_burst} TFLOP/s x {rate:.3g} (MMA rate of this datapath vs kind::f16)"
        if tc_prec in (1, 2) and tf32_meas:
            lib = tf32_meas["burst_tflops"] * (1.0 if tc_prec == 1 else 1.0 / 3.0)
            note += f"; library TF32 GEMM measured in this job: burst {tf32_meas['burst_tflops']}, sustained {tf32_meas['sustained_tflops']} TFLOP/s at {tf32_meas['sm_mhz_sustained']} MHz"
            if lib > peak:
                peak = lib
                note += " (used: larger)"
        roof = {"kernel": f"gemm_conv_tc_kernel ({TC_PRECISIONS[tc_prec]} implicit-GEMM conv/linear)", "bound": "tensor", "achieved": round(achieved, 1),
                "peak": round(peak, 1), "unit": "TFLOP/s", "frac": round(achieved / peak, 3), "traffic": traffic,
                "launches_per_step": tc_n, "flops_per_launch": tc_flops / tc_n, "avg_launch_ms": tc_ms / tc_n,
                "share_of_step": round(tc_ms / total_ms, 3), "peak_basis": note,
                "frac_vs_sustained": round(achieved / (bf16_sust * rate), 3)}
    hbm = {k: {"launches": v[0], "bytes_per_step": int(v[1]), "ms": round(v[2], 3), "achieved": round(v[1] / (v[2] * 1e-3) / 1e9, 1) if v[2] > 0 else None,
               "peak": hbm_gbs, "unit": "GB/s", "frac": round(v[1] / (v[2] * 1e-3) / 1e9 / hbm_gbs, 3) if v[2] > 0 else None}
           for k, v in sorted(fam.items(), key=lambda kv: -kv[1][2])}
    return roof, hbm


def run_ours(args):
    import numpy as np
    import torch
    kind, B, H, W, glob, desc = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    net = build_net(kind, dev, args.precision)
    net.global_motion = glob
    Hp, Wp = pad64(H, W)

    # ---------------- device-resident timing: value ----------------
    g = torch.Generator().manual_seed(1234 + rank)
    im0 = torch.rand(B, 3, Hp, Wp, generator=g).to(dev)
    im1 = torch.rand(B, 3, Hp, Wp, generator=g).to(dev)
    net.zero_copy_outputs = True
    rt = net._runtime
    rt.prepare(net, dev, net.precision, 8, 12)
    plan = rt.plan(B, Hp, Wp, glob)
    launches_per_step = plan.num_launches()
    plan_bytes = getattr(plan, "buffer_bytes", None)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        net(im0, im1)
    t0 = time.time()
    while time.time() - t0 < 1.5:          # let the SM clock ramp; part of warm-up, not timed
        net(im0, im1)
    barrier()
    sampler.reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        net(im0, im1)                      # inputs: 2 x 25 MB per pair at 1080p + ~19 GB of plan buffers >> 126 MB of L2
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * args.steps * B / (ms / 1e3)

    # ---------------- end to end through the reference-facing API: e2e ----------------
    from demo_2x import inference_2frame, interpolate_video
    if args.workload in STREAM_WORKLOADS:
        distinct = synthetic_clip(9, H, W, 7 + rank)                  # 9 distinct frames, replayed back and forth
        order = list(range(9)) + list(range(7, 0, -1))
        clip = lambda n: (distinct[order[i % len(order)]] for i in range(n))
        for _ in interpolate_video(clip(4), net, include_inputs=False):
            pass
        barrier()
        t0 = time.perf_counter()
        n_out = sum(1 for _ in interpolate_video(clip(args.steps + 1), net, include_inputs=False))   # steps pairs -> steps new frames
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        assert n_out == args.steps
        pairs = []
    else:
        pairs = synthetic_u8(B, H, W, 99 + rank)
        if B == 1:      # the step's inputs sit in PINNED host memory (the driver contract): a decoder would write them there
            p0, p1 = net.pinned_frame_buffers(H, W)
            p0[...], p1[...] = pairs[0]
            pairs = [(p0, p1)]
    for a, b in pairs[:1]:
        for _ in range(3):
            inference_2frame(a, b, net)
    if pairs:
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            for a, b in pairs:
                out = inference_2frame(a, b, net)          # pinned staging, H2D, kernels, D2H, sync: all inside
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e = {"value": world * args.steps * B / e2e_s, "unit": "frames/s", "h2d_bytes_per_step": 2 * B * H * W * 3, "d2h_bytes_per_step": B * H * W * 3,
           "api": "demo_2x.inference_2frame (uint8 HWC frames in pinned host memory in, fresh uint8 host frame out)"}
    if args.workload in STREAM_WORKLOADS:
        e2e.update({"h2d_bytes_per_step": B * H * W * 3, "api": "demo_2x.interpolate_video (uint8 HWC host frames in, uint8 host frames out; every frame "
                    "uploaded once, copies overlapped with the neighbouring pair's compute on a second stream)"})
    if B > 1:
        e2e["note"] = "inference_2frame is a batch-1 API: the B pairs of a step are interpolated one after the other"

    # ---------------- N > 1: BASELINE configs[3] in the same invocation - ONE 4K pair in row slabs over all N GPUs -------------
    spatial = None
    if world > 1 and not args.no_spatial:
        rt._plans.clear()
        del plan
        torch.cuda.empty_cache()
        try:
            spatial = spatial_measure(args, "base_4k", net if kind == "base" else build_net("base", dev, args.precision), steps=args.steps, parity=True)
        except Exception as e:          # the pair-sharded line must survive a failure of the slab leg; the failure is reported, not hidden
            spatial = {"error": f"{type(e).__name__}: {e}"[:400]}
        plan = None
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # ---------------- roofline of the dominant kernel + HBM-bound families (rank 0) ----------------
    if plan is None:
        plan = rt.plan(B, Hp, Wp, glob)
    hbm_gbs, bf16_burst, bf16_sust, basis = load_peaks()
    per, tc, fam = measure_kernels(plan, torch)         # BEFORE the 2 s library GEMM below: that one drives the GPU into its power cap
    # ---------------- the other datapaths on the same workload (N = 1): device-resident value + roofline of their GEMM kernel ----------------
    extra = {}
    if world == 1 and not args.no_extra:
        for prec in ("f16", "fp32x3"):
            if prec == args.precision:
                continue
            try:
                extra[prec] = measure_precision(kind, B, Hp, Wp, glob, dev, prec, args.steps, (hbm_gbs, bf16_burst, bf16_sust, basis, None), args.workload)
            except Exception as e:
                extra[prec] = {"error": f"{type(e).__name__}: {e}"[:300]}
    tf32_meas = measure_matmul_peak(torch, dev, "tf32", local) if args.precision in ("tf32", "fp32x3") else None
    total_ms = sum(v[1] for v in per.values())
    roof, roof_hbm = roofline_records(per, tc, fam, args.workload, (hbm_gbs, bf16_burst, bf16_sust, basis, tf32_meas), total_ms)
    kernels = {k: {"launches": v[0], "ms": round(v[1], 3)} for k, v in sorted(per.items(), key=lambda kv: -kv[1][1])}

    # ---------------- comparators (N = 1 only): the reference forward through PyTorch's library kernels on this GPU, and on the host cores ------
    cpu = lib_base = None
    if world == 1 and not args.no_cpu:
        rt._plans.clear()
        del plan
        torch.cuda.empty_cache()
        lib_base = library_baseline(kind, B, Hp, Wp, glob, dev)
        cpu = cpu_baseline(kind, B, Hp, Wp, glob, steps=1)
    line = {
        "metric": "interpolated frames/sec", "value": round(value, 3), "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.precision,
        "data": "synthetic", "config": {"workload": f"{args.workload}: {desc}", "pairs_per_step_per_gpu": B, "padded_shape": [Hp, Wp],
                                        "parallelism": f"pairs sharded over {world} GPU(s), no data-path collective", "weights": "random-init",
                                        "l2": "inputs+activations per step (>1 GB) exceed the 126 MB L2; no flush needed"},
        "e2e": e2e, "gpu_launches": launches_per_step * args.steps, "clocks": clocks, "roofline": roof, "roofline_hbm": roof_hbm,
        "kernels_ms_per_step": kernels, "plan_buffer_bytes": plan_bytes,
    }
    if tf32_meas:
        line["tf32_library_peak"] = tf32_meas
    if cpu:
        line["cpu_baseline"] = cpu
    if lib_base:
        line["gpu_library_baseline"] = lib_base
    if spatial is not None:
        line["spatial_4k"] = spatial
    if extra:
        line["other_precisions"] = extra
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def measure_precision(kind, B, Hp, Wp, glob, dev, precision, steps, peaks, workload):
    """Device-resident frames/s of the same workload on another datapath ("f16": fp16 feature maps + kind::f16 MMAs; "fp32x3": 3xTF32,
    the fp32-tolerance tensor-core path), with the roofline of its GEMM kernel.  Tolerances of each mode: tests/test_gpu_forward.py,
    tests/test_gpu_f16.py."""
    import torch
    net = build_net(kind, dev, precision)
    net.global_motion = glob
    net.zero_copy_outputs = True
    g = torch.Generator().manual_seed(1234)
    im0 = torch.rand(B, 3, Hp, Wp, generator=g).to(dev)
    im1 = torch.rand(B, 3, Hp, Wp, generator=g).to(dev)
    for _ in range(3):
        net(im0, im1)
    t0 = time.time()
    while time.time() - t0 < 1.0:
        net(im0, im1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        net(im0, im1)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    plan = net._runtime.plan(B, Hp, Wp, glob)
    per, tc, fam = measure_kernels(plan, torch)
    total_ms = sum(v[1] for v in per.values())
    roof, roof_hbm = roofline_records(per, tc, fam, workload, peaks, total_ms)
    rec = {"value": round(steps * B / (ms / 1e3), 3), "unit": "frames/s", "ms_per_step": round(ms / steps, 3), "steps": steps, "dtype": precision,
           "roofline": roof, "roofline_hbm": roof_hbm, "plan_buffer_bytes": getattr(plan, "buffer_bytes", None),
           "kernels_ms_per_step": {k: {"launches": v[0], "ms": round(v[1], 3)} for k, v in sorted(per.items(), key=lambda kv: -kv[1][1])}}
    net._runtime._plans.clear()
    del net, plan
    torch.cuda.empty_cache()
    return rec


def spatial_measure(args, workload, net, steps, parity):
    """ONE frame pair per step split into row slabs over the N ranks (NVLink P2P halo exchange, atmvfi/slab.py + atmvfi/p2p.py).
    Strong scaling: value = pairs/s of the whole job; rank 0 receives the interpolated frame.  With `parity`, rank 0 first runs the
    same pair through the single-GPU forward and the slab result must equal it BIT FOR BIT (a row window changes which tiles are
    walked, never the arithmetic of an element).  Every rank calls this; rank 0 returns the record, the others None."""
    import torch
    import torch.distributed as dist
    from atmvfi.p2p import SlabSession
    kind, B, H, W, glob, desc = WORKLOADS[workload]
    rank, world, local = dist.get_rank(), dist.get_world_size(), int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device(f"cuda:{local}")
    net.global_motion = glob
    Hp, Wp = pad64(H, W)
    g = torch.Generator().manual_seed(1234)
    im0 = torch.rand(B, 3, Hp, Wp, generator=g).to(dev)
    im1 = torch.rand(B, 3, Hp, Wp, generator=g).to(dev)
    single = None
    if parity and rank == 0:
        single = net(im0, im1)["I_t"].clone()
        net._runtime._plans.clear()
        torch.cuda.empty_cache()
    sess = SlabSession(net, B, Hp, Wp, gather="I_t")
    launches_per_step = sess.plan.num_launches()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()

    def barrier():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    sess.plan.im0.copy_(im0); sess.plan.im1.copy_(im1)
    for _ in range(max(args.warmup, 3)):
        out = sess.run_inplace(check=False)
    parity_err = None
    if single is not None:
        torch.cuda.synchronize()
        parity_err = float((out["I_t"] - single).abs().max().item())
        del single
    for _ in range(20):                    # clock ramp; a fixed count keeps the ranks' step numbers equal
        sess.run_inplace(check=False)      # pipelined: the peer-time-out word is checked once after the loop
    barrier()
    sampler.reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        sess.run_inplace(check=False)      # pipelined: the peer-time-out word is checked once after the loop
    e1.record()
    barrier()
    sess.check()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    clocks = sampler.stop() if rank == 0 else None
    value = steps * B / (ms / 1e3)

    # e2e: uint8 host frames in pinned memory -> device, slab forward, rank 0 downloads the uint8 result
    a, b = sess.pinned_frame_buffers(H, W)          # inputs in pinned host memory (the driver contract)
    a[...], b[...] = synthetic_u8(1, H, W, 99)[0]
    for _ in range(3):
        sess.interpolate_u8(a, b)
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        sess.interpolate_u8(a, b)
    torch.cuda.synchronize()
    t = torch.tensor([time.perf_counter() - t0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    hb = torch.tensor([float(getattr(sess, "h2d_bytes_this_rank", 2 * H * W * 3))], device=dev)
    dist.all_reduce(hb, op=dist.ReduceOp.SUM)
    h2d = int(hb.item())
    e2e = {"value": steps / float(t.item()), "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": H * W * 3,
           "api": "atmvfi.p2p.SlabSession.interpolate_u8 (inference_2frame arithmetic; uint8 frames in pinned host memory: every rank uploads the rows of its "
                  "slab, the converted rows are all-gathered over NVLink; rank 0 downloads the frame)"}
    barrier()
    per, tc, fam = measure_kernels(sess.plan, torch)      # all ranks replay in lockstep (exchange sites need the peers)
    barrier()
    sess.check()
    st = sess.slab.stats
    rec = None
    if rank == 0:
        hbm_gbs, bf16_burst, bf16_sust, basis = load_peaks()
        total_ms = sum(v[1] for v in per.values())
        roof, roof_hbm = roofline_records(per, tc, fam, workload, (hbm_gbs, bf16_burst, bf16_sust, basis, None), total_ms)
        if roof:
            roof["kernel"] += ", rank 0's row slab"
        kernels = {k: {"launches": v[0], "ms": round(v[1], 3)} for k, v in sorted(per.items(), key=lambda kv: -kv[1][1])}
        rec = {"metric": "interpolated frames/sec", "value": round(value, 3), "unit": "frames/s", "n_gpus": world, "steps": steps, "warmup": max(args.warmup, 3),
               "ms_per_step": round(ms / steps, 3), "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
               "config": {"workload": f"{workload}: {desc}", "pairs_per_step": B, "padded_shape": [Hp, Wp],
                          "parallelism": f"ONE pair per step in {world} row slab(s) {sess.slab.bounds}, halo rows pushed over NVLink P2P",
                          "weights": "random-init", "l2": "activations per step exceed the 126 MB L2; no flush needed"},
               "sites": st["sites"], "pushed_bytes": st["pushed_bytes"], "received_bytes": st["received_bytes"],
               "parity_max_abs_vs_single": parity_err,
               "e2e": e2e, "gpu_launches": launches_per_step * steps, "clocks": clocks, "roofline": roof, "roofline_hbm": roof_hbm, "kernels_ms_per_step": kernels}
    sess.close()
    return rec


def run_spatial(args):
    """--spatial: the row-slab leg on its own (any workload), printed as the bench line."""
    import torch
    import torch.distributed as dist
    kind = WORKLOADS[args.workload][0]
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if not dist.is_initialized():
        if world == 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29655")
            os.environ.setdefault("RANK", "0"); os.environ.setdefault("WORLD_SIZE", "1")
        dist.init_process_group("nccl", device_id=dev)
    net = build_net(kind, dev, args.precision)
    rec = spatial_measure(args, args.workload, net, args.steps, parity=not args.no_parity)
    if rank == 0:
        print(json.dumps(rec))
    dist.destroy_process_group()


def library_baseline(kind, B, Hp, Wp, glob, dev):
    """The reference forward executed by PyTorch's own library kernels (cuDNN / cuBLAS / ATen, eager) on THIS GPU: the oracle (a
    port of the reference's module graph to functional torch ops, oracle/atmvfi_oracle.py) with weights and frames on the device,
    allow_tf32 off and on.  The reference ships no CUDA code of its own, so this is the only "reference on Blackwell" that exists
    (SURVEY 2.1).  A reported comparator next to cpu_baseline; never on the product path."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import atmvfi_oracle as oracle
    import weights
    P = {k: v.to(dev) for k, v in weights.make_weights(kind, "default").items()}
    im0, im1 = [t.to(dev) for t in weights.synthetic_frames(B, Hp, Wp)]
    out = {"api": "oracle.forward (functional torch port of network_base.Network.forward) on cuda: eager cuDNN/cuBLAS/ATen kernels", "unit": "frames/s"}
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark)
    try:
        torch.backends.cudnn.benchmark = True           # as demo_2x.py:18 sets it
        for tag, cudnn_tf32, mm_tf32 in (("fp32", False, False), ("torch_default_cudnn_tf32", True, False), ("all_tf32", True, True)):
            torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = cudnn_tf32, mm_tf32
            with torch.no_grad(), torch.device(dev):
                for _ in range(2):
                    oracle.forward(P, im0, im1, glob)
                torch.cuda.synchronize()
                n = 3
                t0 = time.perf_counter()
                for _ in range(n):
                    oracle.forward(P, im0, im1, glob)
                torch.cuda.synchronize()
                out[tag] = round(n * B / (time.perf_counter() - t0), 3)
    except Exception as e:
        out["error"] = f"{type(e).__name__}: {e}"[:300]
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark = old
    del P, im0, im1
    torch.cuda.empty_cache()
    return out


def cpu_baseline(kind, B, Hp, Wp, glob, steps=1):
    """Times the CPU oracle (oracle/atmvfi_oracle.py, a port of the reference forward) on a bounded sample."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import atmvfi_oracle as oracle
    import weights
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    # bounded sample: one pair of the workload's shape; on a small host (< 16 cores) and >= 1080p, a half-resolution
    # pair of the same kind whose time is scaled by the pixel ratio (the forward is linear in pixels)
    scale, h, w = 1, Hp, Wp
    if cores < 16 and Hp * Wp >= 1088 * 1920:
        h, w, scale = Hp // 2 // 64 * 64, Wp // 2 // 64 * 64, None
        scale = (Hp * Wp) / float(h * w)
    b = 1
    P = weights.make_weights(kind, "default")
    im0, im1 = weights.synthetic_frames(b, h, w)
    t0 = time.perf_counter()
    for _ in range(steps):
        oracle.forward(P, im0, im1, glob)
    dt = (time.perf_counter() - t0) / steps
    pairs_per_s = b / (dt * scale)
    return {"value": round(pairs_per_s, 5), "unit": "frames/s", "cores": cores, "kind": "port",
            "sample": f"1 pair at {h}x{w} (of the {Hp}x{Wp} workload{', time scaled x%.2f by pixel count' % scale if scale > 1 else ''}), fp32, torch CPU threads={cores}, {dt:.1f} s"}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path = the CPU oracle (port), all host threads."""
    kind, B, H, W, glob, desc = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    Hp, Wp = pad64(H, W)
    timed = max(1, min(args.steps, 2))        # bounded sample: each forward is several seconds of all host cores
    cpu = cpu_baseline(kind, B, Hp, Wp, glob, steps=timed)
    line = {"impl": "reference", "metric": "interpolated frames/sec", "value": cpu["value"], "unit": "frames/s", "n_gpus": world, "steps": timed,
            "steps_requested": args.steps, "warmup": 0, "ms_per_step": round(1e3 * B / cpu["value"], 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": f"{args.workload}: {desc}", "padded_shape": [Hp, Wp]},
            "cpu_baseline": cpu, "e2e": {"value": cpu["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="base_1080p", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default="tf32", choices=["tf32", "fp32", "fp32x3", "f16"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--spatial", action="store_true", help="one pair per step split into row slabs over the N GPUs (NVLink P2P halo exchange)")
    ap.add_argument("--no-extra", action="store_true", help="N = 1: skip the extra precision records (f16, fp32x3) next to the headline line")
    ap.add_argument("--no-spatial", action="store_true", help="N > 1: skip the embedded 4K row-slab leg (spatial_4k)")
    ap.add_argument("--no-parity", action="store_true", help="--spatial: skip the slab-vs-single-GPU bit-exactness check")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        if not os.path.exists(os.path.join(ROOT, "atm-vfi_b200", "atmvfi", "libatmvfi_b200.so")):
            sys.path.insert(0, ROOT)
            import __graft_entry__ as ge
            ge.build()
        (run_spatial if args.spatial else run_ours)(args)


if __name__ == "__main__":
    main()
