"""Row-slab scheduling (atmvfi/slab.py) on CPU: N ranks as threads, each with its own buffer set behind the operator
contract emulation; rows move through tests/slab_transports.py.  Unwritten buffer regions are NaN in the emulation,
so a missing halo row or a missing exchange shows up as NaN / a mismatch against the single-rank plan."""
import threading

import pytest
import torch

import weights
from atmvfi.arch import ARCHS
from atmvfi.engine import PackedModel, Plan
from atmvfi.slab import SlabOps, rs_and, rs_sub, rs_union, slab_bounds, win_rows_to_tokens
from atmvfi.ops import WinGeom
from emul_ops import EmulOps
from slab_transports import ThreadTransport, ThreadWorld


def test_interval_sets():
    assert rs_union([(0, 4)], [(4, 8), (10, 12)]) == [(0, 8), (10, 12)]
    assert rs_sub([(0, 10)], [(2, 4), (6, 7)]) == [(0, 2), (4, 6), (7, 10)]
    assert rs_and([(0, 5), (8, 12)], [(3, 9)]) == [(3, 5), (8, 9)]
    assert slab_bounds(2176, 16, 8) == [0, 256, 512, 768, 1088, 1344, 1600, 1856, 2176]
    assert slab_bounds(128, 16, 4) == [0, 32, 64, 96, 128]


def test_window_rows_cover_grid_once():
    for g in (WinGeom(2, 17, 9, 8, 4), WinGeom(2, 136, 16, 12, 6), WinGeom(2, 16, 16, 8, 0)):
        seen = []
        for k in range(g.Hp // g.ws):
            for lo, hi in win_rows_to_tokens(g, k, k + 1):
                seen += list(range(lo, hi))
        assert sorted(seen) == list(range(g.H))


def run_slabs(kind, variant, B, H, W, glob, world, head_major=False):
    P = weights.make_weights(kind, variant)
    im0, im1 = weights.synthetic_frames(B, H, W, kind="texture")
    model = PackedModel(ARCHS[kind], P, 8, 12, with_global=glob)
    ref = Plan(EmulOps(head_major), model, B, H, W, glob).run(im0, im1)
    tw = ThreadWorld(world)
    outs, errs, stats = [None] * world, [], [None] * world

    def worker(r):
        try:
            ops = SlabOps(EmulOps(head_major), r, world, ThreadTransport(tw, r), gather="all")
            plan = Plan(ops, model, B, H, W, glob)
            outs[r] = plan.run(im0, im1)
            stats[r] = ops.stats
        except Exception as e:      # noqa: BLE001
            errs.append((r, e))
            tw.barrier.abort()

    torch.set_num_threads(2)
    ts = [threading.Thread(target=worker, args=(r,)) for r in range(world)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs, errs
    return ref, outs, stats


CASES = [
    ("lite", "stress", 1, 128, 192, True, 2),     # 64-row blocks; padded + shifted global windows straddle the boundary
    ("lite", "stress", 1, 128, 192, True, 4),     # 32-row slabs: every local window straddles
    ("base", "stress", 2, 64, 96, True, 2),       # B = 2
    ("lite", "default", 1, 192, 64, False, 3),    # global off, uneven split
    ("lite", "stress", 1, 256, 64, True, 8),      # 8 ranks x 32 rows: most ranks own no row of 12x12 windows at all, every 8x8 window straddles
]


@pytest.mark.parametrize("kind,variant,B,H,W,glob,world", CASES)
def test_slabs_match_single_rank(kind, variant, B, H, W, glob, world):
    ref, outs, stats = run_slabs(kind, variant, B, H, W, glob, world)
    got = outs[0]
    for key, v in ref.items():
        a = v if isinstance(v, list) else [v]
        b = got[key] if isinstance(v, list) else [got[key]]
        for x, y in zip(a, b):
            assert not torch.isnan(y).any(), key
            assert (x - y).abs().max().item() <= 1e-5, key
    assert all(s["sites"] > 0 for s in stats)


def test_slabs_with_head_major_qkv():
    ref, outs, _ = run_slabs("lite", "stress", 1, 128, 192, True, 4, head_major=True)
    for key in ("I_t", "opt_flow_0", "occ_mask1"):
        assert not torch.isnan(outs[0][key]).any() and (ref[key] - outs[0][key]).abs().max().item() <= 1e-5, key
