"""Parity table: CUDA path vs golden reference outputs for every committed case and precision (run on the GPU box)."""
import glob, json, os, sys
sys.path[:0] = ['atm-vfi_b200', 'atm-vfi_b200/network', 'oracle']
import numpy as np, torch
import weights
def psnr(a, b):
    mse = float(((a.astype(np.float64) - b.astype(np.float64)) ** 2).mean())
    return 99.0 if mse == 0 else -10 * np.log10(mse)
for path in sorted(glob.glob('tests/golden/case_*.npz')):
    if 'config0' in path: continue
    z = np.load(path); meta = json.loads(str(z['meta']))
    P = weights.make_weights(meta['kind'], meta['variant'])
    im0, im1 = weights.synthetic_frames(meta['B'], meta['H'], meta['W'], kind=meta['frames'])
    if meta['kind'] == 'base': from network_base import Network
    else: from network_lite import Network
    net = Network(); net.load_state_dict(P); net = net.cuda().eval(); net.global_motion = meta['global_motion']
    for prec in ('fp32', 'tf32'):
        net.precision = prec
        out = net(im0.cuda(), im1.cuda())
        e = {k: float(np.abs(out[k].cpu().numpy() - z[k]).max()) for k in ('I_t', 'opt_flow_0', 'opt_flow_1', 'occ_mask1', 'I_t_0')}
        lv = [float(np.abs(out['im_t_list'][i].cpu().numpy() - z[f'im_t_list_{i}']).max()) for i in range(len(out['im_t_list']))]
        print(f"{os.path.basename(path)[5:-4]:32s} {prec}: I_t {e['I_t']:.2e} (mean {float(np.abs(out['I_t'].cpu().numpy()-z['I_t']).mean()):.2e}, PSNR {psnr(out['I_t'].cpu().numpy(), z['I_t']):.1f} dB) flow0 {e['opt_flow_0']:.2e} flow1 {e['opt_flow_1']:.2e} occ {e['occ_mask1']:.2e} I_t_0 {e['I_t_0']:.2e} levels {' '.join(f'{v:.1e}' for v in lv)}", flush=True)
