"""A compiled C++ consumer of include/lanczos_b200.h.

oracle/Makefile builds oracle/_ref/abi_consumer_*: the reference's own lanczos_expected (full_TB.h:29-96, read
where it lies in the reference tree) and INTEGRATION.md's `lanczos_expected_b200` binding in ONE translation unit,
linked against liblanczos_b200.so.  The executable fills the reference's static planar arrays, runs both, and
returns memcmp(img_out_ex, img_out_b200) != 0.  Built in the build container (the GPU box has no reference tree;
the binaries travel with the snapshot like the other oracle/_ref/ files).
"""
import glob
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "..", "oracle", "_ref")


def consumers():
    return sorted(glob.glob(os.path.join(REF_DIR, "abi_consumer_*")))


def test_consumers_are_built_and_link_against_the_product_library():
    if not os.path.isdir("/root/reference") and not consumers():
        pytest.skip("no reference tree and no prebuilt consumers")
    exes = consumers()
    assert len(exes) >= 3, "run `make -C oracle` after building the library"
    for exe in exes:
        out = subprocess.run(["ldd", exe], capture_output=True, text=True).stdout
        assert "liblanczos_b200.so" in out and "not found" not in out.split("liblanczos_b200.so")[1].split("\n")[0], out


@pytest.mark.gpu
@pytest.mark.parametrize("content", ["noise", "dark"])
def test_compiled_consumer_matches_the_reference_function(content):
    exes = consumers()
    assert exes, "oracle/_ref/abi_consumer_* missing: build() was not run in the build container"
    for exe in exes:
        res = subprocess.run([exe, content], capture_output=True, text=True, timeout=300)
        assert res.returncode == 0, (exe, res.stdout, res.stderr)
        assert " 0 of " in res.stdout, res.stdout
