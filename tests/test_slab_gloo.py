"""world_size-2 (and 3) run of the row-slab plan over torch.distributed/gloo on CPU: every rank is its own process with its
own buffers; halo rows travel as gloo send/recv of exactly the pieces the scheduler (atmvfi/slab.py) emits."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _worker(rank, world, port, q):
    for p in (os.path.join(HERE, "..", "oracle"), os.path.join(HERE, "..", "atm-vfi_b200"), HERE):
        sys.path.insert(0, os.path.abspath(p))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    import weights
    from atmvfi.arch import ARCHS
    from atmvfi.engine import PackedModel, Plan
    from atmvfi.slab import SlabOps
    from emul_ops import EmulOps
    from slab_transports import GlooTransport
    kind, variant, B, H, W, glob = "lite", "stress", 1, 128, 192, True
    P = weights.make_weights(kind, variant)
    im0, im1 = weights.synthetic_frames(B, H, W, kind="texture")
    model = PackedModel(ARCHS[kind], P, 8, 12, with_global=glob)
    ops = SlabOps(EmulOps(), rank, world, GlooTransport(rank, world), gather="all")
    out = Plan(ops, model, B, H, W, glob).run(im0, im1)
    if rank == 0:
        ref = Plan(EmulOps(), model, B, H, W, glob).run(im0, im1)
        worst = 0.0
        for key, v in ref.items():
            a = v if isinstance(v, list) else [v]
            b = out[key] if isinstance(v, list) else [out[key]]
            for x, y in zip(a, b):
                d = (x - y).abs().max().item()
                worst = max(worst, d if d == d else float("inf"))
        q.put((worst, ops.stats["sites"]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_slab_plan_over_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    worst, sites = q.get(timeout=600)
    [p.join(timeout=120) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert worst <= 1e-5 and sites > 10
