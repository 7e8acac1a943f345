"""One eager forward (no CUDA graph) for ncu captures.  usage: python tools/one_forward.py [base|lite] [H W] [tf32|f16|fp32x3]   (ncu: --profile-from-start off)"""
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
net.precision, net.use_cuda_graph, net.zero_copy_outputs = (sys.argv[4] if len(sys.argv) > 4 else 'tf32'), False, True
im0, im1 = [t.cuda() for t in weights.synthetic_frames(1, H, W)]
net(im0, im1)                      # warm-up: plan build, weight packing
torch.cuda.synchronize()
torch.cuda.profiler.start()        # ncu --profile-from-start off: only the second forward is captured
net(im0, im1)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
