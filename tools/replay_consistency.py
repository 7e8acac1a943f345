"""Race detector of last resort: replay the CUDA graph of one forward many times and check that every output is bit-identical to the
first replay (programmatic dependent launch, buffer recycling and the fused kernels leave no run-to-run difference).
usage: python tools/replay_consistency.py [tf32|f16|fp32x3] [replays]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'atm-vfi_b200'), os.path.join(ROOT, 'atm-vfi_b200', 'network'), os.path.join(ROOT, 'oracle')]
import torch
import weights
from network_base import Network
prec = sys.argv[1] if len(sys.argv) > 1 else 'tf32'
n = int(sys.argv[2]) if len(sys.argv) > 2 else 200
net = Network(); net.load_state_dict(weights.make_weights('base', 'stress')); net = net.cuda().eval(); net.precision = prec
im0, im1 = [t.cuda() for t in weights.synthetic_frames(1, 1088, 1920, kind='texture')]
keys = ('I_t', 'opt_flow_0', 'opt_flow_1', 'occ_mask1')
first, bad = None, 0
for i in range(n):
    out = net(im0, im1)
    torch.cuda.synchronize()
    cur = {k: out[k].clone() for k in keys}
    if first is None:
        first = cur
        assert all(torch.isfinite(v).all() for v in cur.values())
    else:
        bad += sum(0 if torch.equal(first[k], cur[k]) else 1 for k in keys)
print(f"{prec}: {n} replays of the Base 1080p graph, {bad} output tensors differed from the first replay")
sys.exit(1 if bad else 0)
