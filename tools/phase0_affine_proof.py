# phase0_affine_proof.py -- the reference's phase-0 double sum (full_TB.h:58-63 / :71-75 at a coordinate that falls on an
# input sample, a = 3) restated in fp32, and the proof by enumeration that the restatement is exact.
#
# Reference: sum = ((((b0*w0 + b1*w1) + v*1.0) + b3*w3) + b4*w4) + b5*w5 in double, ascending taps, then double_to_uint8
# (truncate, clamp).  w0 = w4 = -alpha, w1 = w3 = +2 alpha (alpha = 1.6e-17, sin(k pi) residues), w5 < 1e-30 (never changes
# the sum).  The result is v or v - 1, and it is decided by how the residues round on the grid of doubles around v
# (spacing u = 2^(e-52) above v in [2^e, 2^(e+1)), u/2 below a power of two).
#
# fp32 has the same grid around v scaled by exactly 2^29 for every binade (24 instead of 53 significand bits), including
# the half spacing below powers of two and the parity of the last bit (v is an integer < 256, so v/u and v/U are even).
# Hence with W_k = fl32(w_k * 2^29):
#     t  = fl32(b0 * W0)
#     y1 = fma32(b1, W1, t)
#     X2 = fl32(v + y1)
#     X3 = fma32(b3, W3, X2)
#     X4 = fma32(b4, W4, X3)
# X_i - v = 2^29 (s_i - v) at every step PROVIDED no fp32 rounding error (W_k carry 24 bits, y1 is rounded twice) moves a
# value across a rounding boundary of the next step.  That is what is enumerated here, stage by stage, over every
# reachable state: (v, b0, b1) -> k2; (v, k2, b3) -> k3; (v, k3, b4) -> k4, where s_i = v + k_i * u/2.
# Then trunc(X4) = trunc(s4) (both in (v - 1, v + 1), same side of v), also for v = 0 (both truncate to 0).
#
# python tools/phase0_affine_proof.py  -> prints the state counts and the smallest distance to a rounding boundary per stage.
import math
import sys

import numpy as np

LD = np.longdouble
F = np.float32


def sinc(x):
    return 1.0 if x == 0 else math.sin(x) / x


def weights(a=3):
    # full_TB.h:16-27 at integer distances a-1-k, k = 0..2a-1
    return np.array([sinc(math.pi * x) * sinc(math.pi * x / a) for x in [float(a - 1 - k) for k in range(2 * a)]])


def fma32(a, b, c):
    """fl32(a*b + c) with a single rounding: the product of two floats and the sum fit a 64-bit significand here."""
    return (a.astype(LD) * b.astype(LD) + c.astype(LD)).astype(F)


def chain32(b0, b1, v, b3, b4, W):
    t = (b0.astype(F) * W[0]).astype(F)
    y1 = fma32(b1.astype(F), np.full(b1.shape, W[1], F), t)
    X2 = (v.astype(F) + y1).astype(F)
    X3 = fma32(b3.astype(F), np.full(b3.shape, W[3], F), X2)
    X4 = fma32(b4.astype(F), np.full(b4.shape, W[4], F), X3)
    return X2, X3, X4


def chain64(b0, b1, v, b3, b4, w):
    s = b0.astype(np.float64) * w[0]
    s = s + b1.astype(np.float64) * w[1]
    s2 = s + v.astype(np.float64) * w[2]
    s3 = s2 + b3.astype(np.float64) * w[3]
    s4 = s3 + b4.astype(np.float64) * w[4]
    return s2, s3, s4


def prove(verbose=True, centres=range(1, 256)):
    w = weights()
    assert w[2] == 1.0 and w[0] < 0 and w[4] < 0 and w[1] > 0 and w[3] > 0 and abs(w[5]) < 1e-30
    W = (w * 2.0 ** 29).astype(F)
    b = np.arange(256)
    B0, B1 = [x.ravel() for x in np.meshgrid(b, b, indexing="ij")]
    total_states = 0
    worst = [1.0, 1.0, 1.0]
    for v in centres:
        e = int(math.floor(math.log2(v)))
        u = 2.0 ** (e - 52)          # spacing of doubles in v's binade; states are kept in units of u/2
        U = u * 2.0 ** 29
        vv = np.full(B0.shape, v)
        # stage 1: (b0, b1) -> k2
        s2, _, _ = chain64(B0, B1, vv, B0, B0, w)
        X2, _, _ = chain32(B0, B1, vv, B0, B0, W)
        k64 = (s2 - v) / (u / 2)
        k32 = (X2.astype(np.float64) - v) / (U / 2)
        if not np.array_equal(k64, k32) or not np.array_equal(k64, np.rint(k64)):
            return False, "stage 1 differs for v = %d" % v
        # distance of the exact real sum to the nearest rounding boundary, in units of the local spacing
        y = (B0 * w[0] + B1 * w[1]) / u
        sp = np.where((y < 0) & (v == 2 ** e), 0.5, 1.0)
        fr = np.abs(y / sp - np.floor(y / sp) - 0.5)
        worst[0] = min(worst[0], fr.min())
        states = np.unique(k64)
        # stage 2: (k2, b3) -> k3
        K, B3 = [x.ravel() for x in np.meshgrid(states, b, indexing="ij")]
        s_in = v + K * (u / 2)
        X_in = (v + K * (U / 2)).astype(F)
        assert np.array_equal(X_in.astype(np.float64), v + K * (U / 2))
        s3 = s_in + B3.astype(np.float64) * w[3]
        X3 = fma32(B3.astype(F), np.full(B3.shape, W[3], F), X_in)
        k64 = (s3 - v) / (u / 2)
        k32 = (X3.astype(np.float64) - v) / (U / 2)
        if not np.array_equal(k64, k32) or not np.array_equal(k64, np.rint(k64)):
            return False, "stage 2 differs for v = %d" % v
        y = K / 2 + B3 * w[3] / u
        sp = np.where((y < 0) & (v == 2 ** e), 0.5, 1.0)
        fr = np.abs(y / sp - np.floor(y / sp) - 0.5)
        worst[1] = min(worst[1], fr.min())
        states3 = np.unique(k64)
        # stage 3: (k3, b4) -> k4
        K, B4 = [x.ravel() for x in np.meshgrid(states3, b, indexing="ij")]
        s_in = v + K * (u / 2)
        X_in = (v + K * (U / 2)).astype(F)
        s4 = s_in + B4.astype(np.float64) * w[4]
        X4 = fma32(B4.astype(F), np.full(B4.shape, W[4], F), X_in)
        k64 = (s4 - v) / (u / 2)
        k32 = (X4.astype(np.float64) - v) / (U / 2)
        if not np.array_equal(k64, k32) or not np.array_equal(k64, np.rint(k64)):
            return False, "stage 3 differs for v = %d" % v
        if not np.array_equal(np.trunc(s4), np.trunc(X4.astype(np.float64))):
            return False, "truncation differs for v = %d" % v
        y = K / 2 + B4 * w[4] / u
        sp = np.where((y < 0) & (v == 2 ** e), 0.5, 1.0)
        fr = np.abs(y / sp - np.floor(y / sp) - 0.5)
        worst[2] = min(worst[2], fr.min())
        total_states += len(states) + len(states3)
        if verbose and (v & (v - 1)) == 0:
            print("v = %3d: %d states after the centre tap (%g .. %g half-spacings), %d after tap 3, flips for %d of %d (k3, b4)"
                  % (v, len(states), states.min(), states.max(), len(states3), int((s4 < v).sum()), len(s4)))
    return True, "all centres agree; %d states; smallest distance to a rounding boundary per stage (in spacings): %.2e %.2e %.2e" % (
        total_states, worst[0], worst[1], worst[2])


def spot_check(n=2_000_000, seed=5):
    """the whole chain on random tuples in several distributions against the reference sum with all six taps"""
    w = weights()
    W = (w * 2.0 ** 29).astype(F)
    rng = np.random.default_rng(seed)
    bad = 0
    for kind in ("uniform", "dark", "lowv", "v0"):
        t = rng.integers(0, 256, size=(6, n))
        if kind == "dark":
            t = rng.integers(0, 24, size=(6, n))
        if kind == "lowv":
            t[2] = rng.integers(0, 12, size=n)
        if kind == "v0":
            t[2] = 0
        s = np.zeros(n)
        for k in range(6):
            s = s + t[k].astype(np.float64) * w[k]
        want = np.clip(np.trunc(s), 0, 255)
        _, _, X4 = chain32(t[0], t[1], t[2], t[3], t[4], W)
        got = np.clip(np.trunc(X4.astype(np.float64)), 0, 255)
        bad += int((want != got).sum())
    return bad


if __name__ == "__main__":
    ok, msg = prove()
    print(msg)
    print("random tuples that differ:", spot_check())
    sys.exit(0 if ok else 1)
