"""Parity report: the CUDA path through the C ABI against the CPU oracle at the BASELINE configurations,
per configuration and content: max |diff|, % exact, % within 1, count above 1 (must be 0), stated separately for the rows
the reference's in-place column pass aliases (rows < K, full_TB.h:67-77) and for the rest, in exact mode and with
LANCZOS_FLAG_TOLERANCE_1LSB.  Test infrastructure (it calls the oracle); needs a GPU.
  python tests/parity_report.py > profiles/<round>_parity_report.txt"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import oracle_py as O  # noqa: E402
import lanczos_hls_b200 as lz  # noqa: E402
from util import dark_hwc, noise_hwc, smooth_hwc  # noqa: E402

CONFIGS = [  # name, in_w, in_h, n, d, a, c
    ("C1 960x540 -> 1920x1080 RGB8 2x", 960, 540, 2, 1, 3, 3),
    ("C2 1920x1080 -> 3840x2160 RGB8 2x", 1920, 1080, 2, 1, 3, 3),
    ("C3 2560x1440 -> 3840x2160 RGBA8 3/2", 2560, 1440, 3, 2, 3, 4),
    ("C4 3840x2160 -> 7680x4320 RGB8 2x (one frame)", 3840, 2160, 2, 1, 3, 3),
    ("C5 shape at 2000^2 -> 3400^2 RGB8 17/10", 2000, 2000, 17, 10, 3, 3),
    ("author's sample 162x89 -> 486x267 RGB8 3x a=2 (host API, widened)", 162, 89, 3, 1, 2, 3),
]


def stats(got, want):
    d = np.abs(got.astype(np.int16) - want.astype(np.int16))
    n = d.size
    return int(d.max()) if n else 0, 100.0 * float((d == 0).sum()) / max(n, 1), 100.0 * float((d <= 1).sum()) / max(n, 1), int((d > 1).sum())


def main():
    print("%-68s %-8s %-9s %3s | %-31s | %-31s" % ("configuration", "content", "mode", "K", "rows < K: max %exact %<=1 n>1", "rows >= K: max %exact %<=1 n>1"))
    worst = 0
    for name, iw, ih, n, d, a, c in CONFIGS:
        ow, oh = O.out_dims(iw, ih, n, d)
        desc = lz.make_desc(iw, ih, ow, oh, c, a, n, d)
        K = lz.alias_rows(desc)
        for kind, gen in (("noise", noise_hwc), ("image", smooth_hwc), ("dark", dark_hwc)):
            img = gen(O, ih, iw, c, seed=7)
            want = O.upscale(img, ow, oh, a, n, d, variant=O.VERBATIM)
            for mode, flags in (("exact", 0), ("1-LSB", lz.FLAG_TOLERANCE_1LSB)):
                if (iw * c) % 4:      # odd width: the host API widens it (device buffers would take the generic kernel)
                    got = lz.upscale(img, ow, oh, a=a, scale_n=n, scale_d=d, flags=flags)
                else:
                    d_in = torch.from_numpy(img).cuda()
                    d_out = torch.empty((oh, ow, c), dtype=torch.uint8, device="cuda")
                    lz.upscale_device(d_in, d_out, a=a, scale_n=n, scale_d=d, flags=flags)
                    torch.cuda.synchronize()
                    got = d_out.cpu().numpy()
                kid = lz.stats()["kernel_id"]
                top, rest = stats(got[:K], want[:K]), stats(got[K:], want[K:])
                print("%-68s %-8s %-9s %3d | %3d %10.6f %10.6f %5d | %3d %10.6f %10.6f %5d   kernel %d" % (
                    name, kind, mode, K, top[0], top[1], top[2], top[3], rest[0], rest[1], rest[2], rest[3], kid))
                worst = max(worst, top[3], rest[3], (top[0] or rest[0]) if mode == "exact" else 0)
    print("RESULT:", "every exact-mode byte identical, every 1-LSB-mode byte within 1" if worst == 0 else "MISMATCH")
    return 1 if worst else 0


if __name__ == "__main__":
    sys.exit(main())
