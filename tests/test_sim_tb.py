"""tools/sim_tb.py: the reference's test bench flow (full_TB.h:99-180: load a picture, run lanczos() and
lanczos_expected(), print the RMS between them, write both pictures) on top of the library."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from util import smooth_hwc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", [(162, 89, 3, 2), (96, 54, 2, 3)], ids=lambda c: "%dx%d_x%d_a%d" % c)
def test_sim_tb_flow(tmp_path, oracle, cfg):
    """The reference template's own size (lanczos.h:13-28: 162x89, 3x, LANCZOS_A 2) and a 2x/a=3 case: PNG in, RMS line,
    two PNGs out with the reference's naming scheme; "expected" is the software path bit for bit, "observed" the
    fixed-point path bit for bit."""
    from PIL import Image
    w, h, scale, a = cfg
    img = smooth_hwc(oracle, h, w, 3, seed=3)
    src = tmp_path / "in.png"
    Image.fromarray(img).save(src)
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sim_tb.py"), str(src), "--scale", str(scale), "--a", str(a),
                          "--out-dir", str(tmp_path)], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr
    assert f"Scale:{scale}/1, WIDTHS {w} -> {w * scale}" in res.stdout          # full_TB.h:124
    rms = float(re.search(r"RMS err: ([0-9.]+)", res.stdout).group(1))          # full_TB.h:166
    stem = f"{w}x{h}->{w * scale}x{h * scale}_{scale}|1_{a}-"                    # full_TB.h:170
    ex = np.asarray(Image.open(tmp_path / (stem + "expected.png")))
    ob = np.asarray(Image.open(tmp_path / (stem + "observed.png")))
    assert np.array_equal(ex, oracle.upscale(img, w * scale, h * scale, a, scale, 1))
    assert np.array_equal(ob, oracle.hls_upscale(img, scale, a, 8))
    want_rms = np.sqrt(((ex.astype(np.int64) - ob.astype(np.int64)) ** 2).sum() / ex.size)
    assert abs(rms - want_rms) < 1e-3 and rms < 12.0, (rms, want_rms)   # ~8: the noisy test content and the de-ring clamp of the HLS path
