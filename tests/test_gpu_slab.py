"""Row windows of every kernel + the slab scheduler on ONE B200: N ranks as threads sharing the device and its default
stream (host enqueue order = execution order), rows moved by device-to-device copies (tests/slab_transports.py).  The
slab result must equal the plain single-rank plan BIT FOR BIT: a row window changes which tiles a kernel walks, never
the arithmetic of an output element.  The multi-GPU NVLink path is exercised by tests/slab_p2p_worker.py (torchrun)."""
import os
import subprocess
import sys
import threading

import pytest
import torch

pytestmark = pytest.mark.gpu

import weights
from atmvfi import _lib
from atmvfi.arch import ARCHS
from atmvfi.engine import PackedModel, Plan, clone_outputs
from atmvfi.ops import CudaOps
from atmvfi.slab import SlabOps
from slab_transports import ThreadTransport, ThreadWorld

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(kind, variant, B, H, W, glob, world, precision):
    dev = torch.device("cuda:0")
    P = {k: v.to(dev) for k, v in weights.make_weights(kind, variant).items()}
    im0, im1 = [t.to(dev) for t in weights.synthetic_frames(B, H, W, kind="texture")]
    model = PackedModel(ARCHS[kind], P, 8, 12, with_global=glob)
    ref = clone_outputs(Plan(CudaOps(dev, precision), model, B, H, W, glob).run(im0, im1))
    tw = ThreadWorld(world)
    outs, errs = [None] * world, []

    def worker(r):
        try:
            torch.cuda.set_device(dev)
            ops = SlabOps(CudaOps(dev, precision), r, world, ThreadTransport(tw, r), gather="all")
            plan = Plan(ops, model, B, H, W, glob)
            outs[r] = plan.run(im0, im1)          # eager: host callables cannot be captured in a graph
            torch.cuda.synchronize()
        except Exception as e:      # noqa: BLE001
            errs.append((r, repr(e)))
            tw.barrier.abort()

    ts = [threading.Thread(target=worker, args=(r,)) for r in range(world)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs, errs
    return ref, outs[0]


CASES = [
    ("lite", "stress", 1, 128, 192, True, 2),
    ("lite", "stress", 1, 128, 192, True, 4),
    ("base", "stress", 2, 64, 96, True, 2),
    ("lite", "default", 1, 192, 64, False, 3),
    ("base", "default", 1, 256, 448, True, 4),
]


@pytest.mark.parametrize("precision", [_lib.FP32, _lib.TF32], ids=["fp32", "tf32"])
@pytest.mark.parametrize("kind,variant,B,H,W,glob,world", CASES)
def test_slabs_bit_exact_on_one_device(kind, variant, B, H, W, glob, world, precision):
    ref, got = _run(kind, variant, B, H, W, glob, world, precision)
    for key, v in ref.items():
        a = v if isinstance(v, list) else [v]
        b = got[key] if isinstance(v, list) else [got[key]]
        for x, y in zip(a, b):
            assert torch.equal(x, y), (key, (x - y).abs().max().item())


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs (gpurun --gpus 2)")
def test_slabs_over_nvlink_p2p():
    n = min(torch.cuda.device_count(), 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29631", os.path.join(ROOT, "tests", "slab_p2p_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "SLAB_P2P_OK" in r.stdout
