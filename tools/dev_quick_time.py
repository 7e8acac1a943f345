import sys, time, torch, subprocess, threading
sys.path[:0] = ['atm-vfi_b200', 'atm-vfi_b200/network', 'oracle']
import weights
from network_base import Network as NB
from network_lite import Network as NL
prec = sys.argv[1] if len(sys.argv) > 1 else 'fp32'
H, W = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (1088, 1920)
kind = sys.argv[4] if len(sys.argv) > 4 else 'base'
B = int(sys.argv[5]) if len(sys.argv) > 5 else 1
net = (NB if kind == 'base' else NL)(); net.load_state_dict(weights.make_weights(kind)); net = net.cuda().eval(); net.precision = prec
im0, im1 = weights.synthetic_frames(B, H, W)
im0, im1 = im0.cuda(), im1.cuda()
net.zero_copy_outputs = True
net.use_cuda_graph = True
clk = []
def sampler():
    p = subprocess.Popen(['nvidia-smi', '--query-gpu=clocks.sm,power.draw', '--format=csv,noheader,nounits', '-lms', '100'], stdout=subprocess.PIPE, text=True)
    sampler.p = p
    for line in p.stdout: clk.append(line.strip())
th = threading.Thread(target=sampler, daemon=True); th.start()
for _ in range(3): net(im0, im1)
torch.cuda.synchronize()
t0 = time.time()
while time.time() - t0 < 2.0: net(im0, im1)      # ramp clocks
torch.cuda.synchronize(); clk.clear()
n = 30
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(n): net(im0, im1)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
sampler.p.kill()
mhz = sorted(int(c.split(',')[0]) for c in clk if c)
print(f'{kind} {prec} B{B} {H}x{W}: {ms:.2f} ms/step  ({B/ms*1e3:.1f} pairs/s)  sm clock median {mhz[len(mhz)//2] if mhz else -1} MHz  samples {clk[:3]}', flush=True)
print('mem GB', torch.cuda.max_memory_allocated()/1e9)
