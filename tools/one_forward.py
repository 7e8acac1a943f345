"""One eager forward (no CUDA graph) for ncu captures.  usage: python tools/one_forward.py [base|lite] [H W]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'atm-vfi_b200'), os.path.join(ROOT, 'atm-vfi_b200', 'network'), os.path.join(ROOT, 'oracle')]
import torch
import weights
from network_base import Network as NB
from network_lite import Network as NL
kind = sys.argv[1] if len(sys.argv) > 1 else 'base'
H, W = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (1088, 1920)
net = (NB if kind == 'base' else NL)(); net.load_state_dict(weights.make_weights(kind)); net = net.cuda().eval()
net.precision, net.use_cuda_graph, net.zero_copy_outputs = 'tf32', False, True
im0, im1 = [t.cuda() for t in weights.synthetic_frames(1, H, W)]
for _ in range(2):
    net(im0, im1)
torch.cuda.synchronize()
print("done")
