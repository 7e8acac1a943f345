"""Drop-in for the reference's benchmark/davis-vid.py (4x demo video from a DAVIS sequence, davis-vid.py:43-136): same command
line (``--TTA``, ``--model_checkpoints``, ``--path``, ``--id``), frames taken two apart, centre crop 480x832, every pair expanded to
frame0, t=0.25, t=0.5, t=0.75 by recursive 2x interpolation.  The three forwards of a pair chain on the device in fp32
(``Network.interpolate_recursive``); the reference moves nothing through uint8 between levels either.

    python benchmark/davis_vid.py --path /data/DAVIS/JPEGImages/480p/ --id breakdance-flare --model_checkpoints ckpt.pt
"""
import argparse
import glob
import os
import os.path as osp
import sys

import numpy as np

_HERE = osp.dirname(osp.abspath(__file__))
for _p in (osp.dirname(_HERE), osp.join(osp.dirname(_HERE), "network")):
    if _p not in sys.path:
        sys.path.insert(0, _p)


def interpolate_sequence(model, frames_bgr, time_interval=2, H=480, W=832, interpolate4x=True, TTA=False):
    """frames_bgr: list of HxWx3 uint8 BGR arrays -> generator of output frames in display order (davis-vid.py:89-135)."""
    last = None
    for i in range(0, len(frames_bgr) - time_interval, time_interval):
        f0, f1 = frames_bgr[i], frames_bgr[i + time_interval]
        H_, W_, _ = f0.shape
        crop = lambda f: np.ascontiguousarray(f[H_ // 2 - H // 2: H_ // 2 + H // 2, W_ // 2 - W // 2: W_ // 2 + W // 2])
        mids = model.interpolate_recursive_u8(crop(f0), crop(f1), levels=2 if interpolate4x else 1, isBGR=True, divisor=64, TTA=TTA)
        yield f0                    # the reference writes the UN-cropped input frame here (davis-vid.py:120)
        for m in mids:
            yield m
        last = f1
    if last is not None:
        yield last


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--TTA", type=lambda s: str(s).lower() in ("1", "true", "yes"), default=False)
    ap.add_argument("--model_checkpoints", type=str, default="")
    ap.add_argument("--path", type=str, default="./DAVIS/JPEGImages/480p/")
    ap.add_argument("--id", type=str, default="breakdance-flare")
    ap.add_argument("--model_type", choices=["base", "lite"], default="base")
    ap.add_argument("--out", type=str, default="./video/output.mp4")
    args = ap.parse_args()
    import cv2
    from demo_2x import load_model_checkpoint
    from network_base import Network as NB
    from network_lite import Network as NL
    model = (NB if args.model_type == "base" else NL)()
    if args.model_checkpoints:
        load_model_checkpoint(model, args.model_checkpoints)
    model = model.to("cuda").eval()
    files = sorted(glob.glob(osp.join(args.path, args.id, "*.jpg")))
    frames = [cv2.imread(f) for f in files]
    os.makedirs(osp.dirname(osp.abspath(args.out)), exist_ok=True)
    out = cv2.VideoWriter(args.out, cv2.VideoWriter_fourcc(*"mp4v"), 10.0, (832, 480))
    for f in interpolate_sequence(model, frames, TTA=args.TTA):
        out.write(f)
    out.release()


if __name__ == "__main__":
    main()
