"""torchrun worker: one frame pair split into row slabs over the visible GPUs through CUDA IPC + NVLink peer stores
(atmvfi/p2p.py, csrc/p2p.cu); rank 0 compares the gathered outputs with its own single-GPU plan (must be bit-exact).
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/slab_p2p_worker.py [H W]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "oracle"), os.path.join(ROOT, "atm-vfi_b200"), os.path.join(ROOT, "atm-vfi_b200", "network")):
    sys.path.insert(0, p)

import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import weights
    from atmvfi.engine import clone_outputs
    from atmvfi.p2p import SlabSession
    from network_base import Network as Base
    from network_lite import Network as Lite
    cases = [("lite", "stress", 1, 128, 192, True), ("base", "default", 1, 256, 448, True), ("lite", "default", 2, 192, 128, False)]
    if len(sys.argv) >= 3:
        cases = [("base", "default", 1, int(sys.argv[1]), int(sys.argv[2]), True)]
    ok = True
    for kind, variant, B, H, W, glob in cases:
        for precision in ("fp32", "tf32"):
            P = weights.make_weights(kind, variant)
            net = (Base if kind == "base" else Lite)()
            net.load_state_dict(P, strict=True)
            net = net.to(f"cuda:{local}").eval()
            net.global_motion, net.precision = glob, precision
            im0, im1 = [t.cuda() for t in weights.synthetic_frames(B, H, W, kind="texture")]
            sess = SlabSession(net, B, H, W, gather="all")
            for use_graph in (False, True, True):
                out = sess.run(im0, im1, use_graph=use_graph)
                sess.check()
                if rank == 0:
                    got = clone_outputs(out)
            dist.barrier()
            st = sess.slab.stats
            if rank == 0:
                net.zero_copy_outputs = False
                ref = net(im0, im1)
                worst = 0.0
                for key, v in ref.items():
                    a = v if isinstance(v, list) else [v]
                    b = got[key] if isinstance(v, list) else [got[key]]
                    for x, y in zip(a, b):
                        worst = max(worst, (x - y).abs().max().item())
                        if not torch.equal(x, y):
                            ok = False
                            print(f"MISMATCH {kind} {H}x{W} {precision} {key}: max|d| = {(x - y).abs().max().item():.3e}", flush=True)
                print(f"[slab p2p] {kind} {variant} B={B} {H}x{W} glob={glob} {precision} world={world}: max|slab - single| = {worst:.3e}; "
                      f"sites={st['sites']} pushed={st['pushed_bytes'] / 1e6:.1f} MB received={st['received_bytes'] / 1e6:.1f} MB "
                      f"launches={sess.plan.num_launches()}", flush=True)
            sess.close()
            dist.barrier()
    # uint8 front end: every rank uploads only the rows of its slab, the planar rows are all-gathered over NVLink; two different
    # pairs in a row (the second one must not see rows of the first) against the single-GPU inference_2frame arithmetic
    import numpy as np
    for kind, H, W in (("lite", 200, 300), ("base", 250, 180)):
        P = weights.make_weights(kind, "default")
        net = (Base if kind == "base" else Lite)()
        net.load_state_dict(P, strict=True)
        net = net.to(f"cuda:{local}").eval()
        rng = np.random.default_rng(7)                         # same frames on every rank
        frames = [rng.integers(0, 256, (H, W, 3), dtype=np.uint8) for _ in range(3)]
        Hp, Wp = H + (-H) % 64, W + (-W) % 64
        sess = SlabSession(net, 1, Hp, Wp, gather="I_t")
        for a, b in ((frames[0], frames[1]), (frames[1], frames[2]), (frames[2], frames[0])):
            got = sess.interpolate_u8(a, b, copy=True)
            if rank == 0:
                want = net.interpolate_u8(a, b)
                same = np.array_equal(got, want)
                ok = ok and same
                print(f"[slab p2p u8] {kind} {H}x{W} world={world} sliced upload: {'identical' if same else 'MISMATCH'} "
                      f"(this rank uploads {sess.h2d_bytes_this_rank} of {2 * H * W * 3} bytes)", flush=True)
        sess.close()
        dist.barrier()
    if rank == 0:
        print("SLAB_P2P_OK" if ok else "SLAB_P2P_FAILED", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
