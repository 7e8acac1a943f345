import sys, torch
sys.path[:0] = ['atm-vfi_b200', 'atm-vfi_b200/network', 'oracle']
import weights
from network_base import Network as NB
from network_lite import Network as NL
kind = sys.argv[1]; prec = sys.argv[2]; H, W = int(sys.argv[3]), int(sys.argv[4]); B = int(sys.argv[5]) if len(sys.argv) > 5 else 1
net = (NB if kind == 'base' else NL)(); net.load_state_dict(weights.make_weights(kind)); net = net.cuda().eval(); net.precision = prec
net.use_cuda_graph = False; net.zero_copy_outputs = True
im0, im1 = weights.synthetic_frames(B, H, W); im0, im1 = im0.cuda(), im1.cuda()
for _ in range(2): net(im0, im1)
torch.cuda.synchronize()
