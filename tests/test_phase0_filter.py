"""CPU proof-by-enumeration of the slow paths' phase-0 re-check (lanczos_v6.cu phase0_doubt2, plan.cpp).

For an output coordinate exactly on an input sample the reference's weights are 1 at the centre tap and
sin(k*pi) residues (~1e-17) elsewhere (full_TB.h:39-53 has no |x| < a window), so its double sum
(full_TB.h:58-63) truncates to v or to v-1.  The kernels copy v and use a conservative test to find the
samples that might be v-1; only those are re-evaluated in double.  This file restates that test with the
library's own constants (lanczos_b200_phase0_constants) in exact arithmetic -- an fp16 FMA is emulated as
a float64 FMA-free expression, exact for fp16 operands, followed by ONE rounding to float16 -- and checks
on millions of tap tuples that a clear doubt bit really implies "the reference returns v".
"""
import math

import numpy as np
import pytest


def ref_weights(a=3):
    """L(x) at the integer offsets of a phase-0 sample, as full_TB.h:39-53 computes them (libm sin)."""
    def sinc(x):
        return 1.0 if x == 0 else math.sin(x) / x
    return np.array([sinc(math.pi * x) * sinc(math.pi * x / a) for x in [float(a - 1 - k) for k in range(2 * a)]])


def ref_phase0(b, w):
    """full_TB.h:58-63 + double_to_uint8 (:29-37) on tap tuples b[n][6]: plain double multiply, then add."""
    s = np.zeros(b.shape[0], dtype=np.float64)
    for k in range(6):
        s = s + b[:, k].astype(np.float64) * w[k]        # numpy: one rounding per operation, no FMA
    return np.clip(np.trunc(s), 0, 255).astype(np.int64)


def h16(bits):
    return np.array([bits & 0xFFFF], dtype=np.uint16).view(np.float16).astype(np.float64)[0]


def fma16(x, y, z):
    """fp16 FMA: operands are fp16 values held in float64, the exact x*y+z fits float64, one rounding."""
    return (x * y + z).astype(np.float16).astype(np.float64)


def doubt2(b, consts):
    """phase0_doubt2 of lanczos_v6.cu, one fp16 lane, vectorised over tuples. True = "may be v-1"."""
    nk0, k1, k3, nk4 = (h16(c) for c in consts)
    sub = 2.0 ** -24                                     # bytes arrive as fp16 subnormals b * 2^-24
    b0, b1, v, b3, b4 = (b[:, k].astype(np.float64) * sub for k in range(5))
    v12 = (v * 4096.0).astype(np.float16).astype(np.float64)
    t = fma16(v12, 2.0, -(2.0 ** -12))
    H = (t.astype(np.float16).view(np.uint16) & 0x7C00).view(np.float16).astype(np.float64)
    nH2 = (H * -2.0).astype(np.float16).astype(np.float64)
    zpre = fma16(b1, k1, fma16(b0, nk0, H))
    z1 = fma16(b4, nk4, H)
    z2 = fma16(b3, k3, fma16(b4, nk4, nH2))
    neg = lambda z: np.signbit(z.astype(np.float16))
    return neg(zpre) | (neg(z1) & neg(z2))


def tuples(rng, n, kind):
    b = rng.integers(0, 256, size=(n, 6), dtype=np.int64)
    if kind == "dark":
        b &= 15
    elif kind == "dark_centre":
        b[:, 2] &= 31
    elif kind == "pow2_centre":
        b[:, 2] = 1 << rng.integers(0, 8, size=n)
    elif kind == "bright_neighbours":
        b[:, [0, 4]] |= 0xC0
        b[:, 2] &= 63
    return b


@pytest.mark.parametrize("ratio", [(2, 1), (3, 2)])
def test_clear_doubt_bit_implies_reference_returns_v(lz, ratio):
    n_, d_ = ratio
    consts = lz.phase0_constants(lz.make_desc(8 * d_, 8 * d_, 8 * n_, 8 * n_, 3, 3, n_, d_))
    assert all(c != 0 for c in consts) and consts[0] >> 31 and consts[3] >> 31 and not consts[1] >> 31
    w = ref_weights(3)
    rng = np.random.default_rng(20261018)
    n_flip = n_doubt = n_total = 0
    for kind in ("uniform", "dark", "dark_centre", "pow2_centre", "bright_neighbours"):
        b = tuples(rng, 400_000, kind)
        ref = ref_phase0(b, w)
        flip = ref != b[:, 2]
        assert np.all((ref == b[:, 2]) | (ref == b[:, 2] - 1))       # v or v-1, nothing else
        doubt = doubt2(b, consts) & (b[:, 2] != 0)                      # v = 0 is skipped by the kernels: clamps to 0
        assert not np.any(flip & ~doubt), f"{kind}: a sample the test calls safe is v-1 in the reference"
        n_flip += int(flip.sum()); n_doubt += int(doubt.sum()); n_total += b.shape[0]
    # the test must stay useful: it may over-report, but not by orders of magnitude
    assert n_flip > 0 and n_doubt < 8 * n_flip + n_total // 50


def test_exhaustive_small_centres(lz):
    """Every (v, b0, b1, b3, b4) with v <= 8 and neighbours on a grid that includes all small values."""
    consts = lz.phase0_constants(lz.make_desc(8, 8, 16, 16, 3, 3, 2, 1))
    w = ref_weights(3)
    grid = np.array(sorted(set(list(range(0, 24)) + list(range(24, 256, 15)) + [255])), dtype=np.int64)
    g0, g1, g3, g4 = np.meshgrid(grid, grid, grid, grid, indexing="ij")
    for v in range(0, 9):
        b = np.stack([g0.ravel(), g1.ravel(), np.full(g0.size, v), g3.ravel(), g4.ravel(), np.full(g0.size, 255)], axis=1)
        ref = ref_phase0(b, w)
        doubt = doubt2(b, consts) & (v != 0)
        assert not np.any((ref != v) & ~doubt), f"v={v}"


def test_zero_centre_never_flips():
    """v = 0: the residues sum to something tiny of either sign; the quantiser clamps it to 0 = v."""
    w = ref_weights(3)
    rng = np.random.default_rng(7)
    b = rng.integers(0, 256, size=(200_000, 6), dtype=np.int64)
    b[:, 2] = 0
    assert np.all(ref_phase0(b, w) == 0)
