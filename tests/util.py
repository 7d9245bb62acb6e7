"""Shared helpers for the tests: synthetic images and the golden fixture list."""
import glob
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def golden_files():
    return sorted(glob.glob(os.path.join(HERE, "golden", "ref_*.npz")))


def noise_hwc(oracle, h, w, c, seed=0):
    """Uniform xorshift noise, interleaved [H][W][C]."""
    return oracle.xorshift_bytes(h * w * c, oracle.SEED + seed).reshape(h, w, c)


def smooth_hwc(oracle, h, w, c, seed=0):
    """Smooth + noise: 128 + 90 sin(0.05x+c) cos(0.037y) + U[-8,7]."""
    yy, xx = np.mgrid[0:h, 0:w]
    noise = (oracle.xorshift_bytes(h * w * c, oracle.SEED + 77 + seed).reshape(h, w, c).astype(np.int32) & 15) - 8
    img = np.stack([128 + 90 * np.sin(0.05 * xx + ch) * np.cos(0.037 * yy) for ch in range(c)], axis=-1)
    return np.clip(img + noise, 0, 255).astype(np.uint8)


def dark_hwc(oracle, h, w, c, seed=0):
    """Dark noise 0..15: the regime where the reference's sin(k*pi) residues flip phase-0 samples."""
    return (noise_hwc(oracle, h, w, c, seed) & 15).astype(np.uint8)


def planar(img_hwc):
    return np.ascontiguousarray(np.transpose(img_hwc, (2, 0, 1)))


def interleaved(img_chw):
    return np.ascontiguousarray(np.transpose(img_chw, (1, 2, 0)))


def diff_stats(a, b):
    d = np.abs(a.astype(np.int16) - b.astype(np.int16))
    return {"max": int(d.max()) if d.size else 0, "exact": float((d == 0).mean()) if d.size else 1.0,
            "n_diff": int((d != 0).sum())}
