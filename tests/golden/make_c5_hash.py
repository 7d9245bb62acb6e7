#!/usr/bin/env python
"""Known answer for BASELINE configs[4] at FULL size: the CPU oracle (pinned to the compiled
reference, tests/test_oracle.py) upscales one 16384x16384 RGB8 uniform-noise image x1.7 (17/10) to 27852x27852 and
the FNV-1a-64 of the interleaved output goes to tests/golden/c5_hash.txt.  About 15 minutes on 8 cores; run once in the
build container.  tests/test_gpu_parity.py::test_config5_full_size_hash recomputes the input on the GPU box (same
xorshift seed), runs the row-band API and compares hashes: no 2.3 GB fixture travels.
  python tests/golden/make_c5_hash.py [size]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle_py as O  # noqa: E402

size = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
ow = oh = size * 17 // 10
t0 = time.time()
img = O.xorshift_bytes(size * size * 3, O.SEED + 5).reshape(size, size, 3)
out = O.upscale(img, ow, oh, 3, 17, 10)
h = O.fnv1a64(out)
line = f"{size} {size} {ow} {oh} 17 10 3 3 {h:016x}\n"
path = os.path.join(ROOT, "tests", "golden", "c5_hash.txt")
lines = []
if os.path.exists(path):
    lines = [ln for ln in open(path) if not ln.startswith(f"{size} ")]
if not lines:
    lines = ["# in_w in_h out_w out_h n d a c fnv1a64(oracle output, interleaved y,x,c) ; input = xorshift seed SEED+5 interleaved; tests/golden/make_c5_hash.py\n"]
open(path, "w").writelines(lines + [line])
print(line.strip(), f"({time.time() - t0:.0f} s)")
