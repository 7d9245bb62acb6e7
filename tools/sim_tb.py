#!/usr/bin/env python
"""sim_tb.py -- the reference's full test bench (full_TB.h:99-180) on the B200 library.

Loads an image (PNG/JPEG via Pillow instead of stb_image, full_TB.h:107), runs the fixed-point
"HLS mode" (the `lanczos()` arithmetic, full_TB.h:140) and the software path (`lanczos_expected`,
full_TB.h:141) on the GPU, prints the RMS error between them like full_TB.h:166 and writes the two
images with the reference's naming scheme (full_TB.h:170-177).

    python tools/sim_tb.py in.png --scale 2 --a 3 [--out-dir img/]
Without an input file a synthetic 162x89 image (the size of the reference's params.h template,
lanczos.h:19-20) is used.
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("image", nargs="?")
    ap.add_argument("--scale", type=int, default=2, help="integer SCALE_N (SCALE_D = 1: the HLS path's valid range)")
    ap.add_argument("--a", type=int, default=3, help="LANCZOS_A")
    ap.add_argument("--bit-precision", type=int, default=8, help="BIT_PRECISION (lanczos.h:28)")
    ap.add_argument("--out-dir", default=".")
    args = ap.parse_args()

    import torch
    from PIL import Image
    import lanczos_hls_b200 as lz

    if args.image:
        img = np.asarray(Image.open(args.image).convert("RGB"), dtype=np.uint8)
    else:
        yy, xx = np.mgrid[0:89, 0:162]
        img = np.stack([128 + 90 * np.sin(0.05 * xx + c) * np.cos(0.037 * yy) for c in range(3)], -1).clip(0, 255).astype(np.uint8)
    h, w, c = img.shape
    oh, ow = h * args.scale, w * args.scale
    print("Running full TB")
    print(f"Scale:{args.scale}/1, WIDTHS {w} -> {ow}")                      # full_TB.h:124

    d_in = torch.from_numpy(np.ascontiguousarray(img)).cuda()
    d_ob = torch.empty((oh, ow, c), dtype=torch.uint8, device="cuda")
    d_ex = torch.empty((oh, ow, c), dtype=torch.uint8, device="cuda")
    lz.upscale_hls_device(d_in, d_ob, a=args.a, bit_precision=args.bit_precision)   # "observed"  (lanczos)
    lz.upscale_device(d_in, d_ex, a=args.a, scale_n=args.scale, scale_d=1)          # "expected"  (lanczos_expected)
    torch.cuda.synchronize()
    ob, ex = d_ob.cpu().numpy(), d_ex.cpu().numpy()
    err = ((ex.astype(np.int64) - ob.astype(np.int64)) ** 2).sum()
    print("RMS err: %.3f" % np.sqrt(err / (c * ow * oh)))                  # full_TB.h:166
    stem = f"{w}x{h}->{ow}x{oh}_{args.scale}|1_{args.a}-"                    # full_TB.h:170
    os.makedirs(args.out_dir, exist_ok=True)
    Image.fromarray(ex).save(os.path.join(args.out_dir, stem + "expected.png"))
    Image.fromarray(ob).save(os.path.join(args.out_dir, stem + "observed.png"))
    print("wrote", stem + "expected.png", "and", stem + "observed.png", "to", args.out_dir)


if __name__ == "__main__":
    main()
