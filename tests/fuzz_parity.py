"""Randomised differential test: CUDA path (C ABI) vs the CPU oracle on many small random shapes, ratios,
contents (noise, image-like, dark noise, patchworks of the three) and layouts.  usage: [FUZZ_MAXW=300 FUZZ_MAXH=120] python tests/fuzz_parity.py [seconds] [seed]
(needs a GPU; exits 1 on a mismatch)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import oracle_py as O
import lanczos_hls_b200 as lz
from util import noise_hwc, smooth_hwc, dark_hwc, planar

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 12345)
MAXW, MAXH = int(os.environ.get("FUZZ_MAXW", "300")), int(os.environ.get("FUZZ_MAXH", "120"))
RATIOS = [(2, 1), (2, 1), (3, 2), (17, 10), (3, 1), (5, 3), (4, 1), (1, 1), (7, 4)]
t0, n, kernels = time.time(), 0, {}
while time.time() - t0 < budget:
    nn, dd = RATIOS[rng.integers(len(RATIOS))]
    a = int(rng.choice([3, 3, 3, 2, 1, 4]))
    c = int(rng.choice([3, 3, 4, 1, 2]))
    iw = int(rng.integers(2 * a + 1, MAXW)); ih = int(rng.integers(2 * a + 1, MAXH))
    if rng.random() < 0.6:                       # shapes the specialised kernels accept
        iw = max(16, iw // 16 * 16)
    ow, oh = O.out_dims(iw, ih, nn, dd)
    if ow < 1 or oh < ih:
        continue
    kind = rng.integers(4)
    if kind < 3:
        img = [noise_hwc, smooth_hwc, dark_hwc][kind](O, ih, iw, c, seed=int(rng.integers(1 << 20)))
    else:                                        # patchwork of the three (the phase-0 regimes switch inside a segment)
        sd, band, col = int(rng.integers(1 << 20)), int(rng.integers(3, 60)), int(rng.integers(8, 400))
        parts = [f(O, ih, iw, c, seed=sd + i) for i, f in enumerate((smooth_hwc, noise_hwc, dark_hwc))]
        yy, xx = np.mgrid[0:ih, 0:iw]
        sel = ((yy // band) + (xx // col)) % 3
        img = np.zeros((ih, iw, c), dtype=np.uint8)
        for i in range(3):
            img[sel == i] = parts[i][sel == i]
    flags = int(rng.choice([0, 0, lz.FLAG_NO_ALIAS]))
    variant = O.CLEAN if flags & lz.FLAG_NO_ALIAS else O.VERBATIM
    mode = rng.integers(3)
    if mode == 2:                                # planar entry point
        want = O.expected_planar(planar(img), ow, oh, a, nn, dd, variant=variant, fast=True)
        d_in = torch.from_numpy(planar(img)).cuda(); d_out = torch.empty((c, oh, ow), dtype=torch.uint8, device="cuda")
        lz.upscale_planar_device(d_in, d_out, a=a, scale_n=nn, scale_d=dd, flags=flags)
    else:
        want = O.upscale(img, ow, oh, a, nn, dd, variant=variant)
        d_in = torch.from_numpy(img).cuda(); d_out = torch.empty((oh, ow, c), dtype=torch.uint8, device="cuda")
        if mode == 1 and oh >= 8:                # two row bands with their own halo rows
            desc = lz.make_desc(iw, ih, ow, oh, c, a, nn, dd, flags=flags)
            cut = int(rng.integers(1, oh))
            for r0, r1 in ((0, cut), (cut, oh)):
                lo, cnt = lz.band_input_rows(desc, r0, r1 - r0)
                lz.upscale_band_device(desc, d_in[lo:lo + cnt], d_out[r0:r1], r0, r1 - r0, lo, cnt)
        else:
            lz.upscale_device(d_in, d_out, a=a, scale_n=nn, scale_d=dd, flags=flags)
    torch.cuda.synchronize()
    kid = lz.stats()["kernel_id"]; kernels[kid] = kernels.get(kid, 0) + 1
    got = d_out.cpu().numpy()
    if not np.array_equal(got, want):
        bad = np.argwhere(got != want)
        print(f"MISMATCH iw={iw} ih={ih} {nn}/{dd} a={a} c={c} kind={kind} flags={flags} mode={mode} kernel={kid}: "
              f"{len(bad)} bytes, first at {bad[0].tolist()}, max |diff| {np.abs(got.astype(int) - want.astype(int)).max()}")
        sys.exit(1)
    n += 1
print(f"{n} random cases bit-exact in {time.time() - t0:.0f} s; kernel ids used: {dict(sorted(kernels.items()))}")
