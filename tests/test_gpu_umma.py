"""GPU: the tcgen05 building blocks (descriptor format, shared-memory operand
layout, TMEM accumulator read-back, bf16x3 split) against a float64 product."""
import numpy as np
import pytest
import torch

from gossipnet_b200 import ops

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('ts', [False, True])
@pytest.mark.parametrize('k', [16, 32, 64, 96, 256])
def test_umma_selftest_matches_float64(k, ts):
    rs = np.random.RandomState(k)
    a = rs.normal(0, 1, (128, k)).astype(np.float32)
    w = rs.normal(0, 1, (k, 64)).astype(np.float32)
    got = ops.selftest_umma(torch.from_numpy(a).cuda(), torch.from_numpy(w).cuda(),
                            a_in_tmem=ts).cpu().numpy()
    ref = a.astype(np.float64) @ w.astype(np.float64)
    err = np.max(np.abs(got - ref)) / np.max(np.abs(ref))
    assert err < 3e-5, err      # bf16x3: ~2^-16 relative per product
    # and it really is better than plain bf16 (which would be ~4e-3)
    assert err < 1e-4


@pytest.mark.parametrize('ts', [False, True])
def test_umma_selftest_exact_on_small_integers(ts):
    """Integers < 256 are exact in bf16: the result must be bit-exact, which
    pins the operand layout (any row / k permutation error would show)."""
    rs = np.random.RandomState(0)
    a = rs.randint(-8, 9, (128, 64)).astype(np.float32)
    w = rs.randint(-8, 9, (64, 64)).astype(np.float32)
    got = ops.selftest_umma(torch.from_numpy(a).cuda(), torch.from_numpy(w).cuda(),
                            a_in_tmem=ts).cpu().numpy()
    assert np.array_equal(got, a @ w)
