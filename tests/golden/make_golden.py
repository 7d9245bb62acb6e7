"""Generate golden input/output vectors from the REFERENCE ITSELF.

Runs in the build container only (needs /root/reference): `python tests/golden/make_golden.py`.
It builds oracle/_ref (the reference's full_TB.h:29-96 compiled where it lies, one object per
compile-time configuration, see oracle/Makefile) and stores, per configuration, the synthetic
input and the reference's lanczos_expected() output in tests/golden/<cfg>.npz.
The fixtures are committed; the reference tree is not needed to run the tests.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle_py as O  # noqa: E402

# (in_w, in_h, out_w, out_h, n, d, a, c), input kind
CONFIGS = [
    ((96, 54, 192, 108, 2, 1, 3, 3), "noise"),
    ((96, 54, 192, 108, 2, 1, 2, 3), "noise"),
    ((96, 54, 144, 81, 3, 2, 3, 3), "noise"),
    ((100, 60, 170, 102, 17, 10, 3, 3), "noise"),
    ((64, 48, 96, 72, 3, 2, 3, 4), "smooth"),
    ((54, 30, 162, 90, 3, 1, 2, 3), "noise"),
    ((37, 23, 74, 46, 2, 1, 3, 4), "edges"),
    ((50, 40, 85, 68, 17, 10, 2, 3), "smooth"),
    ((33, 17, 132, 68, 4, 1, 3, 1), "impulse"),
]


def make_input(kind, c, h, w, seed):
    if kind == "noise":
        return O.xorshift_bytes(c * h * w, O.SEED + seed).reshape(c, h, w)
    yy, xx = np.mgrid[0:h, 0:w]
    if kind == "smooth":
        noise = (O.xorshift_bytes(c * h * w, O.SEED + seed).reshape(c, h, w).astype(np.int32) & 15) - 8
        planes = [128 + 90 * np.sin(0.05 * xx + ch) * np.cos(0.037 * yy) for ch in range(c)]
        return np.clip(np.stack(planes) + noise, 0, 255).astype(np.uint8)
    if kind == "edges":
        img = np.zeros((c, h, w), np.uint8)
        img[:, :, w // 2:] = 255
        img[:, h // 2:, :] ^= 255
        img[0] = ((xx + yy) & 1) * 255  # 1-px checkerboard
        return img
    if kind == "impulse":
        img = np.zeros((c, h, w), np.uint8)
        img[:, h // 2, w // 2] = 255
        img[:, 0, 0] = 255
        img[:, h - 1, w - 1] = 200
        return img
    raise ValueError(kind)


def main():
    if not os.path.isdir("/root/reference/LanczosUpscaler"):
        sys.exit("reference tree not present: fixtures can only be regenerated in the build container")
    O.build()
    here = os.path.dirname(os.path.abspath(__file__))
    for i, (cfg, kind) in enumerate(CONFIGS):
        iw, ih, ow, oh, n, d, a, c = cfg
        img = make_input(kind, c, ih, iw, i)
        out = O.ref_expected_planar(img, cfg)
        name = f"ref_{iw}x{ih}_{ow}x{oh}_{n}_{d}_a{a}_c{c}_{kind}.npz"
        np.savez_compressed(os.path.join(here, name), cfg=np.array(cfg, np.int32), img_in=img, img_out=out)
        print(name, hex(O.fnv1a64(out)))
    # the two 960x540 known answers are kept as hashes only (6 MB each otherwise)
    with open(os.path.join(here, "ref_hashes.txt"), "w") as fh:
        fh.write("# in_w in_h out_w out_h n d a c fnv1a64(reference output, planar c,y,x) ; input = xorshift seed SEED planar\n")
        for cfg in [(960, 540, 1920, 1080, 2, 1, 3, 3), (960, 540, 1920, 1080, 2, 1, 2, 3),
                    (1920, 135, 3840, 270, 2, 1, 3, 3)]:
            iw, ih, ow, oh, n, d, a, c = cfg
            img = O.xorshift_bytes(c * ih * iw).reshape(c, ih, iw)
            out = O.ref_expected_planar(img, cfg)
            fh.write(" ".join(str(v) for v in cfg) + f" {O.fnv1a64(out):016x}\n")
            print(cfg, f"{O.fnv1a64(out):016x}")


if __name__ == "__main__":
    main()
