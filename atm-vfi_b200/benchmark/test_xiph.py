"""Drop-in for the reference's benchmark/test_xiph.py (Xiph 2K / "4K" accuracy evaluation, test_xiph.py:56-150): same command line
(``--root``, ``--TTA``, ``--ckpt``), same two categories (frames resized to 2048x1080 with INTER_AREA, centre 2048x1080 crop of the
4096x2160 frames), odd frames in, even frame as ground truth, ``InputPadder(divisor=32)``, ``global_motion = True``, flip TTA.
The reference downloads the clips with ffmpeg; here the PNG frames must already be under ``--root/<clip>/%03d.png``.  Without
them (``--synthetic N``) the same loop runs on N moving-texture triplets of the two categories' shape.

    python benchmark/test_xiph.py --root /data/xiph --ckpt ckpt.pt [--TTA True] [--model_type base|lite]
"""
import argparse
import glob
import os
import os.path as osp
import sys

import numpy as np
import torch

_HERE = osp.dirname(osp.abspath(__file__))
for _p in (osp.dirname(_HERE), osp.join(osp.dirname(_HERE), "network")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

from benchmark.harness import evaluate_triplets, synthetic_triplets      # noqa: E402

FILE_LIST = ['BoxingPractice', 'Crosswalk', 'DrivingPOV', 'FoodMarket', 'FoodMarket2', 'RitualDance', 'SquareAndTimelapse', 'Tango']


def xiph_triplets(root, clip, category):
    import cv2
    read = lambda f: cv2.imread(f)[:, :, ::-1]
    d = osp.join(root, clip)
    for k in range(2, 99, 2):
        frames = [read(f'{d}/{j:03d}.png') for j in (k - 1, k, k + 1)]
        if category == 'resized-2k':
            frames = [cv2.resize(src=f, dsize=(2048, 1080), fx=0.0, fy=0.0, interpolation=cv2.INTER_AREA) for f in frames]
        else:
            frames = [f[540:-540, 1024:-1024, :] for f in frames]
        yield frames[0], frames[1], frames[2]


def run(model, root=None, TTA=False, synthetic=0, log=print):
    model.global_motion = True
    results = {}
    for category in ['resized-2k', 'cropped-4k']:
        ps, ss, n = [], [], 0
        clips = FILE_LIST if not synthetic else ['synthetic']
        for clip in clips:
            if synthetic:
                trip = synthetic_triplets(synthetic, 1080, 2048, seed=len(category))
            else:
                if len(glob.glob(osp.join(root, clip, '*.png'))) < 100:
                    raise SystemExit(f'{osp.join(root, clip)} does not hold the 100 extracted frames (see the reference test_xiph.py:79-98 for the ffmpeg recipe)')
                trip = xiph_triplets(root, clip, category)
            r = evaluate_triplets(model, trip, divisor=32, TTA=TTA)
            ps.append(r["psnr"] * r["n"]); ss.append(r["ssim"] * r["n"]); n += r["n"]
            log(f'[Xiph] [{category}/{clip}] psnr: {sum(ps) / n:.02f}, ssim: {sum(ss) / n:.04f}')
        results[category] = {"psnr": sum(ps) / n, "ssim": sum(ss) / n, "n": n}
    return results


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('-r', '--root', default='./xiph')
    ap.add_argument("--TTA", type=lambda s: str(s).lower() in ("1", "true", "yes"), default=False)
    ap.add_argument("--ckpt", type=str, default="")
    ap.add_argument("--model_type", choices=["base", "lite"], default="base")
    ap.add_argument("--synthetic", type=int, default=0, help="evaluate N synthetic 2048x1080 triplets instead of the Xiph frames")
    args = ap.parse_args()
    from demo_2x import load_model_checkpoint
    from network_base import Network as NB
    from network_lite import Network as NL
    model = (NB if args.model_type == "base" else NL)()
    if args.ckpt:
        load_model_checkpoint(model, args.ckpt)
    model = model.to('cuda').eval()
    print(run(model, args.root, args.TTA, args.synthetic))


if __name__ == "__main__":
    main()
