"""Drop-in for the reference's benchmark/test_snufilm.py (SNU-FILM easy / medium / hard / extreme, test_snufilm.py:76-152): same
command line (``--path``, ``--img_data_path``, ``--TTA``, ``--ckpt``), the four ``test-*.txt`` triplet lists, ``InputPadder(divisor=64)``,
``global_motion = True``, ``ensemble_global_motion = False``, flip TTA, PSNR on [0,1] and the "matlab" SSIM.  ``--synthetic N``
runs the same loop on N moving-texture 1280x720 triplets per level when the dataset is not on disk.

    python benchmark/test_snufilm.py --path /data/snufilm/eval_modes/ --img_data_path /data/snufilm/ --ckpt ckpt.pt
"""
import argparse
import os
import os.path as osp
import sys

_HERE = osp.dirname(osp.abspath(__file__))
for _p in (osp.dirname(_HERE), osp.join(osp.dirname(_HERE), "network")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

from benchmark.harness import evaluate_triplets, synthetic_triplets      # noqa: E402

LEVELS = ['test-easy.txt', 'test-medium.txt', 'test-hard.txt', 'test-extreme.txt']


def snufilm_triplets(path, img_data_path, test_file):
    import cv2
    with open(osp.join(path, test_file), "r") as f:
        for line in f:
            names = line.replace("data/SNU-FILM/test/", img_data_path).strip().split(' ')
            i0, it, i1 = (cv2.imread(osp.join(path, n))[:, :, ::-1] for n in names[:3])
            yield i0, it, i1


def run(model, path=None, img_data_path=None, TTA=False, synthetic=0, log=print):
    model.global_motion = True
    model.ensemble_global_motion = False
    results = {}
    for k, test_file in enumerate(LEVELS):
        trip = synthetic_triplets(synthetic, 720, 1280, seed=k) if synthetic else snufilm_triplets(path, img_data_path, test_file)
        r = evaluate_triplets(model, trip, divisor=64, TTA=TTA)
        log('Testing level:' + test_file[:-4])
        log('Avg PSNR: {} SSIM: {}'.format(r["psnr"], r["ssim"]))
        results[test_file[:-4]] = r
    return results


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--path", type=str, default="./snufilm-test/eval_modes/")
    ap.add_argument("--img_data_path", type=str, default="./snufilm-test/")
    ap.add_argument("--device", type=str, default='cuda')
    ap.add_argument("--TTA", type=lambda s: str(s).lower() in ("1", "true", "yes"), default=False)
    ap.add_argument("--ckpt", type=str, default="")
    ap.add_argument("--model_type", choices=["base", "lite"], default="base")
    ap.add_argument("--synthetic", type=int, default=0)
    args = ap.parse_args()
    from demo_2x import load_model_checkpoint
    from network_base import Network as NB
    from network_lite import Network as NL
    model = (NB if args.model_type == "base" else NL)()
    if args.ckpt:
        load_model_checkpoint(model, args.ckpt)
    model = model.to(args.device).eval()
    print(f'=========================Starting testing=========================')
    print(f'Dataset: SNU_FILM\t     TTA: {args.TTA}')
    run(model, args.path, args.img_data_path, args.TTA, args.synthetic)


if __name__ == "__main__":
    main()
