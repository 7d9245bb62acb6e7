"""Host-side sharding of the hot path across GPUs (one process per GPU, no collectives).

The reference has no process/device boundary; its only scaling device is the
2a-line cyclic buffer plus ROW_WORKERS output rows per block (worker.h:132,
lanczos.cpp:72-81).  The B200 analogue:
  * frame batches: contiguous frame ranges per rank;
  * one large image: contiguous output-row bands per rank, each rank reading its
    own overlapping halo rows (`band_input_rows`), phases from the global row index.
"""


def split_range(total, rank, world):
    """Contiguous [lo, hi) share of `total` items for `rank` of `world` (sizes differ by at most 1)."""
    if world < 1 or not (0 <= rank < world) or total < 0:
        raise ValueError("bad shard arguments")
    return total * rank // world, total * (rank + 1) // world


def frame_range(n_frames, rank, world):
    return split_range(n_frames, rank, world)


def band_range(out_h, rank, world):
    return split_range(out_h, rank, world)


def band_input_rows_py(out_row0, out_rows, in_h, a, scale_n, scale_d, alias_rows=0, alias_in_rows=0):
    """Pure-Python twin of lanczos_b200_band_input_rows (plan.cpp band_rows) for CPU-side tests:
    input rows floor(r0*D/N)-a+1 .. floor((r1-1)*D/N)+a clipped to the image (full_TB.h:72)."""
    lo = min(in_h - 1, max(0, out_row0 * scale_d // scale_n - a + 1))
    hi = min(in_h - 1, (out_row0 + out_rows - 1) * scale_d // scale_n + a)
    if out_row0 < alias_rows:
        lo, hi = 0, max(hi, alias_in_rows - 1)
    hi = max(hi, lo)
    return lo, hi - lo + 1
