"""Replay selected GEMM-shaped launches of a plan on their own (for ncu): one full forward first (so every buffer holds real data),
then each selected launch once.   usage: python tools/run_layers.py <precision> <idx> [<idx> ...]
<idx> counts the atmvfi_gemm_conv records of the Base 1080p plan in launch order (tools/layer_table.py prints them in that order).
Under ncu: -k regex:gemm_conv_tc -s <number of gemm launches in the plan> -c <number of indices>."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'atm-vfi_b200'), os.path.join(ROOT, 'atm-vfi_b200', 'network'), os.path.join(ROOT, 'oracle')]
import torch
import weights
from network_base import Network as NB
prec = sys.argv[1]
idx = [int(a) for a in sys.argv[2:]]
net = NB(); net.load_state_dict(weights.make_weights('base')); net = net.cuda().eval(); net.precision = prec
rt = net._runtime; rt.prepare(net, torch.device('cuda:0'), prec, 8, 12)
plan = rt.plan(1, 1088, 1920, True)
im0, im1 = [t.cuda() for t in weights.synthetic_frames(1, 1088, 1920)]
plan.run(im0, im1, use_graph=False)
torch.cuda.synchronize()
gemms = [r for r in plan.records if r[0] == 'atmvfi_gemm_conv']
print("gemm launches in the plan:", len(gemms))
plan.ops.set_rounding()
st = torch.cuda.current_stream().cuda_stream
for i in idx:
    name, fn, args, keep = gemms[i]
    fn(*args, st)
    torch.cuda.synchronize()
    print(i, keep[2].name)
