import sys, time, torch
sys.path[:0] = ['atm-vfi_b200', 'atm-vfi_b200/network', 'oracle']
import weights
from network_base import Network
prec = sys.argv[1] if len(sys.argv) > 1 else 'fp32'
H, W = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (1088, 1920)
net = Network(); net.load_state_dict(weights.make_weights('base')); net = net.cuda().eval(); net.precision = prec
im0, im1 = weights.synthetic_frames(1, H, W)
im0, im1 = im0.cuda(), im1.cuda()
net.zero_copy_outputs = True
for g in (False, True):
    net.use_cuda_graph = g
    for _ in range(2): net(im0, im1)
    torch.cuda.synchronize(); t = time.time()
    n = 3
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): net(im0, im1)
    e1.record(); torch.cuda.synchronize()
    print(f'{prec} {H}x{W} graph={g}: {e0.elapsed_time(e1)/n:.2f} ms/pair (wall {(time.time()-t)/n*1e3:.2f})', flush=True)
print('mem GB', torch.cuda.max_memory_allocated()/1e9)
