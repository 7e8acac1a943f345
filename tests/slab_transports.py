"""Row-exchange transports for the CPU tests of atmvfi/slab.py - TEST INFRASTRUCTURE ONLY.

The product moves halo rows with NVLink peer stores (atmvfi/p2p.py -> csrc/p2p.cu).  These stand-ins move the same
rows between the per-rank buffer sets of (a) threads of one process and (b) gloo ranks, so that the scheduling logic
(which rows, from whom, before which operator) is exercised without a GPU.
"""
from __future__ import annotations

import threading
from collections import defaultdict

import torch
import torch.distributed as dist


class ThreadWorld:
    def __init__(self, world: int):
        self.world = world
        self.slabs = [None] * world
        self.sems = defaultdict(lambda: threading.Semaphore(0))
        self.barrier = threading.Barrier(world)
        self.lock = threading.Lock()

    def sem(self, key):
        with self.lock:
            return self.sems[key]


class ThreadTransport:
    def __init__(self, tw: ThreadWorld, rank: int):
        self.tw, self.rank = tw, rank

    def attach(self, slab):
        self.slab = slab
        self.tw.slabs[self.rank] = slab

    def step_begin(self):
        self.slab.backend.emit_host(lambda: self.tw.barrier.wait(timeout=120))

    def exchange(self, site, outgoing, incoming):
        tw, me = self.tw, self.rank

        def run():
            for ps in outgoing:
                dst = tw.slabs[ps.dst].bufs[ps.buf.idx]
                a, b = ps.img0 * ps.buf.planes, (ps.img0 + ps.nimg) * ps.buf.planes
                dst.rows3[a:b, ps.lo : ps.hi] = ps.buf.rows3[a:b, ps.lo : ps.hi]
            for d in sorted({ps.dst for ps in outgoing}):
                tw.sem((site, me, d)).release()
            for s in sorted({ps.src for ps in incoming}):
                assert tw.sem((site, s, me)).acquire(timeout=120), f"rank {me}: no rows from rank {s} at site {site}"

        self.slab.backend.emit_host(run)


class GlooTransport:
    """torch.distributed (gloo) send/recv of the same row pieces."""

    def __init__(self, rank: int, world: int):
        self.rank, self.world = rank, world

    def attach(self, slab):
        self.slab = slab

    def step_begin(self):
        self.slab.backend.emit_host(lambda: dist.barrier())

    def exchange(self, site, outgoing, incoming):
        def run():
            reqs, stash = [], []
            for ps in outgoing:
                a, b = ps.img0 * ps.buf.planes, (ps.img0 + ps.nimg) * ps.buf.planes
                reqs.append(dist.isend(ps.buf.rows3[a:b, ps.lo : ps.hi].contiguous(), ps.dst))
            for ps in incoming:
                a, b = ps.img0 * ps.buf.planes, (ps.img0 + ps.nimg) * ps.buf.planes
                tmp = torch.empty_like(ps.buf.rows3[a:b, ps.lo : ps.hi]).contiguous()
                reqs.append(dist.irecv(tmp, ps.src))
                stash.append((ps, a, b, tmp))
            for r in reqs:
                r.wait()
            for ps, a, b, tmp in stash:
                ps.buf.rows3[a:b, ps.lo : ps.hi] = tmp

        self.slab.backend.emit_host(run)
