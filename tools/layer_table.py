"""Per-layer device time of the plan's GEMM-shaped launches (CUDA events), for tuning.  Run on the GPU box:
    python tools/layer_table.py [base|lite] [H W] [B] [tf32|f16|fp32x3]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'atm-vfi_b200'), os.path.join(ROOT, 'atm-vfi_b200', 'network'), os.path.join(ROOT, 'oracle')]
import torch
import weights
kind = sys.argv[1] if len(sys.argv) > 1 else 'base'
H, W = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (1088, 1920)
B = int(sys.argv[4]) if len(sys.argv) > 4 else 1
PREC = sys.argv[5] if len(sys.argv) > 5 else 'tf32'
from network_base import Network as NB
from network_lite import Network as NL
net = (NB if kind == 'base' else NL)(); net.load_state_dict(weights.make_weights(kind)); net = net.cuda().eval(); net.precision = PREC
rt = net._runtime; rt.prepare(net, torch.device('cuda:0'), PREC, 8, 12)
plan = rt.plan(B, H, W, True)
ops = plan.ops; recs = plan.records
ops.set_rounding()
st = torch.cuda.current_stream().cuda_stream
import time
t0 = time.time()
while time.time() - t0 < 1.5: plan.launch()
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(recs) + 1)]
acc = [0.0] * len(recs)
R = 5
for rep in range(R):
    ev[0].record()
    for i, (name, fn, args, _) in enumerate(recs):
        fn(*args, st); ev[i + 1].record()
    torch.cuda.synchronize()
    for i in range(len(recs)): acc[i] += ev[i].elapsed_time(ev[i + 1]) / R
tot = sum(acc)
print(f"total {tot:.2f} ms over {len(recs)} launches")
for i, (name, fn, args, keep) in enumerate(recs):
    if name == 'atmvfi_gemm_conv':
        d, srcs, w = keep[0], keep[1], keep[2]
        cin = sum(s.C for s in srcs); n = d.Cout * (4 if d.out_mode == 1 else 1)
        fl = 2.0 * d.B * d.Hout * d.Wout * cin * d.ksize ** 2 * n
        byt = float(srcs[0].esize) * d.B * d.Hin * d.Win * cin + float(keep[3].esize) * d.B * d.Hout * d.Wout * n
        print(f"{acc[i]*1e3:8.1f} us {fl/acc[i]/1e9:7.1f} TF/s {byt/acc[i]/1e6:7.0f} GB/s  { {0: 'f32', 1: 'tc ', 2: 'x3 ', 3: 'f16'}[d.precision] } k{d.ksize} s{d.stride} {d.B}x{d.Hout}x{d.Wout} cin={cin} cout={d.Cout}{' deconv' if d.out_mode==1 else ''}{' winrev' if d.out_mode==2 else ''}{' +res' if d.residual else ''}  {w.name}")
    else:
        print(f"{acc[i]*1e3:8.1f} us {'':32s} {name}")
