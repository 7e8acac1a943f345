"""Drop-in for the reference's demo_2x.py: ``load_model_checkpoint``, ``inference_2frame`` and the 2x video CLI.

    python demo_2x.py --model_type base --ckpt path/to/ckpt.pt --frame0 a.png --frame1 b.png --out out.png
    python demo_2x.py --model_type lite --ckpt ckpt.pt --video in.mp4
"""
import argparse
import os
import sys
import warnings

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
for _p in (_HERE, os.path.join(_HERE, "network")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

from network_base import Network as Network_base      # noqa: E402
from network_lite import Network as Network_lite      # noqa: E402

warnings.filterwarnings('ignore')
torch.set_grad_enabled(False)


def load_model_checkpoint(model, checkpoint_path, strict=True):
    """Same contract as demo_2x.py:24-51: accepts the trainer's wrapped checkpoint or a raw state-dict,
    drops the lazily registered ``attn_mask`` / ``HW`` buffers, loads with ``strict`` and returns the
    optimizer state (``None`` for a raw state-dict, where the reference raises UnboundLocalError)."""
    print(f'--- loading from checkpoint: {checkpoint_path} ---')
    checkpt = torch.load(checkpoint_path, map_location='cuda:0' if torch.cuda.is_available() else 'cpu')
    optim_checkpt = None
    if isinstance(checkpt, dict) and 'model_state_dict' in checkpt:
        param = checkpt['model_state_dict']
        optim_checkpt = checkpt.get('optimizer_state_dict')
        if 'meta_data' in checkpt:
            print(checkpt['meta_data'])
        print(f"\t- train: {checkpt.get('train_metric')}\n\t- val: {checkpt.get('val_metric')}")
    else:
        param = checkpt
    param = {k: v for k, v in param.items() if 'attn_mask' not in k and 'HW' not in k}
    model.load_state_dict(param, strict=strict)
    n = sum(p.numel() for p in model.parameters() if p.requires_grad)
    print(f"total trainable parameters: {round(n / 1e6, 2)} M")
    return optim_checkpt


def inference_2frame(img0, img1, model, isBGR=True):
    """img0, img1: numpy [H,W,3] uint8 -> numpy [H,W,3] uint8 middle frame (demo_2x.py:54-87)."""
    return model.interpolate_u8(np.ascontiguousarray(img0), np.ascontiguousarray(img1), isBGR=isBGR, divisor=64)


def inference_multiframe(img0, img1, model, levels=2, isBGR=True, TTA=False):
    """2^levels x interpolation of one pair (benchmark/davis-vid.py:102-106 for levels=2): numpy uint8 frames in, the
    2^levels - 1 in-between uint8 frames out; the recursion stays on the device in fp32 (Network.interpolate_recursive)."""
    return model.interpolate_recursive_u8(np.ascontiguousarray(img0), np.ascontiguousarray(img1), levels=levels, isBGR=isBGR, divisor=64, TTA=TTA)


def interpolate_video(frames, model, isBGR=True, include_inputs=True):
    """The 2x loop of demo_2x.py:129-168 over an iterable of uint8 frames: yields f0, mid, f1, mid, ..., f_last.  Same result
    per pair as ``inference_2frame``; uploads, downloads and compute of neighbouring pairs overlap (model.interpolate_stream)."""
    return model.interpolate_stream(frames, isBGR=isBGR, divisor=64, include_inputs=include_inputs)


def _build(model_type, ckpt, global_off):
    model = (Network_base if model_type == 'base' else Network_lite)()
    if ckpt:
        load_model_checkpoint(model, ckpt)
    model = model.to('cuda').eval()
    model.global_motion = not global_off
    return model


def main():
    ap = argparse.ArgumentParser(description='ATM-VFI 2x interpolation on B200')
    ap.add_argument('--model_type', choices=['base', 'lite'], default='base')
    ap.add_argument('--ckpt', default='')
    ap.add_argument('--frame0'); ap.add_argument('--frame1'); ap.add_argument('--out', default='output.png')
    ap.add_argument('--video'); ap.add_argument('--out_video', default='output_2x.mp4')
    ap.add_argument('--global_off', action='store_true')
    args = ap.parse_args()
    import cv2
    model = _build(args.model_type, args.ckpt, args.global_off)
    if args.video:
        cap = cv2.VideoCapture(args.video)
        fps = cap.get(cv2.CAP_PROP_FPS)
        ok, first = cap.read()
        if not ok:
            raise SystemExit(f'cannot read {args.video}')
        h, w = first.shape[:2]
        wr = cv2.VideoWriter(args.out_video, cv2.VideoWriter_fourcc(*'mp4v'), 2 * fps, (w, h))

        def decoded():
            yield first
            while True:
                ok, cur = cap.read()
                if not ok:
                    return
                yield cur

        for frame in interpolate_video(decoded(), model):
            wr.write(frame)
        wr.release(); cap.release()
    else:
        a, b = cv2.imread(args.frame0), cv2.imread(args.frame1)
        cv2.imwrite(args.out, inference_2frame(a, b, model))


if __name__ == '__main__':
    main()
