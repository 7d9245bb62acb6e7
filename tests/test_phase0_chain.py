"""CPU proof-by-enumeration of the phase-0 second look (lanczos_v6.cu phase0_chain2 / v_fix_phase0_chunk / h_fix,
plan.cpp verify_phase0_chain).

For an output coordinate exactly on an input sample the reference's weights are 1 at the centre tap and sin(k*pi)
residues (~1e-17) elsewhere (full_TB.h:39-53 has no |x| < a window), so its double sum (full_TB.h:58-63) truncates to
the centre value v or to v-1.  The kernels copy v, use a cheap conservative filter to find the samples that might be
v-1, and decide those with an fp32 restatement of the sum: the grid of floats around the integer v is the grid of doubles
scaled by 2^29, so the sum with residues scaled by 2^29 rounds the way the reference's does.  Whether the fp32 rounding
of the scaled residues can ever change a decision is settled by enumerating every reachable state of the sum
(tools/phase0_affine_proof.py); this file runs that enumeration with the LIBRARY's constants, checks that the library
ran its own (C++) enumeration successfully, and checks the whole chain on random tap tuples against the reference.
"""
import importlib.util
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_spec = importlib.util.spec_from_file_location("phase0_affine_proof", os.path.join(ROOT, "tools", "phase0_affine_proof.py"))
proof = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(proof)


def ref_phase0(b, w):
    """full_TB.h:58-63 + double_to_uint8 (:29-37) on tap tuples b[n][6]: plain double multiply, then add."""
    s = np.zeros(b.shape[0], dtype=np.float64)
    for k in range(6):
        s = s + b[:, k].astype(np.float64) * w[k]        # numpy: one rounding per operation, no FMA
    return np.clip(np.trunc(s), 0, 255).astype(np.int64)


@pytest.mark.parametrize("ratio", [(2, 1), (3, 2), (3, 1)])
def test_library_verified_the_chain_and_uses_the_proved_constants(lz, ratio):
    n_, d_ = ratio
    ok, consts = lz.phase0_chain(lz.make_desc(8 * d_, 8 * d_, 8 * n_, 8 * n_, 3, 3, n_, d_))
    assert ok, "plan.cpp verify_phase0_chain failed: the specialised kernels would not be used"
    w = proof.weights(3)
    want = (w * 2.0 ** 29).astype(np.float32)
    assert consts[2] == 1.0
    for k in (0, 1, 3, 4):
        assert np.float32(consts[k]) == want[k], (k, consts[k], want[k])


def test_no_chain_for_other_kernel_sizes(lz):
    for a in (1, 2, 4):
        ok, consts = lz.phase0_chain(lz.make_desc(8, 8, 16, 16, 3, a, 2, 1))
        assert not ok


def test_enumeration_of_every_state_of_the_sum():
    """Every (v, b0, b1) -> state after the centre tap; every (v, state, b3); every (v, state, b4): the fp32 chain and
    the reference's doubles agree on the state (in half-spacings around v) at every step and on the truncated result."""
    ok, msg = proof.prove(verbose=False)
    assert ok, msg


def test_whole_chain_on_random_tuples():
    assert proof.spot_check(n=500_000) == 0


def test_result_is_v_or_v_minus_one_and_both_occur():
    w = proof.weights(3)
    rng = np.random.default_rng(20261019)
    b = rng.integers(0, 256, size=(400_000, 6), dtype=np.int64)
    ref = ref_phase0(b, w)
    assert np.all((ref == b[:, 2]) | (ref == b[:, 2] - 1))
    flips = int((ref != b[:, 2]).sum())
    assert 0 < flips < b.shape[0] // 20                     # ~2 % on uniform noise
    W = (w * 2.0 ** 29).astype(np.float32)
    _, _, X4 = proof.chain32(b[:, 0], b[:, 1], b[:, 2], b[:, 3], b[:, 4], W)
    assert np.array_equal(np.clip(np.trunc(X4.astype(np.float64)), 0, 255).astype(np.int64), ref)


def test_zero_centre_never_flips():
    """v = 0: the residues sum to something tiny of either sign; the quantiser clamps it to 0 = v."""
    w = proof.weights(3)
    rng = np.random.default_rng(7)
    b = rng.integers(0, 256, size=(200_000, 6), dtype=np.int64)
    b[:, 2] = 0
    assert np.all(ref_phase0(b, w) == 0)
