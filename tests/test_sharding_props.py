"""Property tests (hypothesis) of the host-side sharding logic: frame ranges and row bands with halo rows
(SURVEY.md 8e).  CPU only: the C ABI's band arithmetic is host code; the pixel-level statement is checked
with the oracle, which implements the same band semantics (phases from the global row index)."""
import numpy as np
from hypothesis import given, settings, strategies as st

from lanczos_hls_b200.sharding import band_input_rows_py, band_range, frame_range, split_range

RATIOS = [(2, 1), (3, 2), (17, 10), (3, 1), (5, 3), (1, 1), (7, 4)]


@given(total=st.integers(0, 5000), world=st.integers(1, 16))
def test_ranges_partition_the_work(total, world):
    edges = [split_range(total, r, world) for r in range(world)]
    assert edges[0][0] == 0 and edges[-1][1] == total
    for (lo, hi), (lo2, _) in zip(edges, edges[1:]):
        assert lo <= hi == lo2
    sizes = [hi - lo for lo, hi in edges]
    assert max(sizes) - min(sizes) <= 1
    assert frame_range(total, 0, world) == edges[0] and band_range(total, world - 1, world) == edges[-1]


@settings(max_examples=60, deadline=None)
@given(in_h=st.integers(8, 400), ratio=st.sampled_from(RATIOS), a=st.integers(1, 4), world=st.integers(1, 8),
       no_alias=st.booleans())
def test_band_rows_match_c_abi_and_cover_every_tap(lz, in_h, ratio, a, world, no_alias):
    n, d = ratio
    out_h = in_h * n // d
    desc = lz.make_desc(16, in_h, 16 * n // d, out_h, 3, a, n, d, flags=lz.FLAG_NO_ALIAS if no_alias else 0)
    k0 = lz.alias_rows(desc)
    # rows the in-place emulation needs: up to the last row read by an aliased row (plan.cpp)
    last = lambda y: min(in_h - 1, y * d // n + a)
    top = max([last(y) for y in range(k0)], default=-1)
    alias_in = max([last(y) for y in range(top + 1)], default=-1) + 1
    for rank in range(min(world, out_h)):
        r0, r1 = band_range(out_h, rank, min(world, out_h))
        if r1 == r0:
            continue
        lo, cnt = lz.band_input_rows(desc, r0, r1 - r0)
        assert (lo, cnt) == band_input_rows_py(r0, r1 - r0, in_h, a, n, d, k0, alias_in)
        # every tap row of every output row of the band (full_TB.h:72) lies inside [lo, lo + cnt)
        for y in (r0, (r0 + r1) // 2, r1 - 1):
            first, lst = max(0, y * d // n - a + 1), min(in_h - 1, y * d // n + a)
            if first <= lst:
                assert lo <= first and lst < lo + cnt


@settings(max_examples=20, deadline=None)
@given(periods=st.integers(6, 14), ratio=st.sampled_from([(2, 1), (3, 2), (17, 10)]), m0=st.integers(2, 4),
       seed=st.integers(0, 1000))
def test_band_from_its_own_halo_rows_equals_the_full_image(oracle, periods, ratio, m0, seed):
    """A band that starts on a ratio period can be computed from a crop of the input that holds its halo rows:
    the rows above the crop only matter to the first output rows of the crop, which lie outside the band."""
    n, d = ratio
    a, in_w = 3, 20
    in_h = d * periods
    out_w, out_h = oracle.out_dims(in_w, in_h, n, d)
    img = oracle.xorshift_bytes(in_h * in_w * 3, oracle.SEED + seed).reshape(in_h, in_w, 3)
    full = oracle.upscale(img, out_w, out_h, a, n, d, variant=oracle.CLEAN)
    halo = -(-(a - 1) // d) + 1                       # ratio periods of input kept above the band
    crop0 = d * (m0 - halo)                           # first input row of the crop (>= 0: m0 >= halo)
    if crop0 < 0:
        return
    crop = img[crop0:]
    c_out_h = oracle.out_dims(in_w, crop.shape[0], n, d)[1]
    part = oracle.upscale(crop, out_w, c_out_h, a, n, d, variant=oracle.CLEAN)
    r0 = n * m0                                       # first output row of the band in the full image
    assert np.array_equal(part[n * halo:], full[r0:r0 + part.shape[0] - n * halo])
