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


def test_tma_conventions():
    """Tensor-map TMA (tile load + gather4) into SWIZZLE_128B tiles and UMMAs on SWIZZLE_128B
    K-major descriptors: the conventions of gn_block_tma.cu, pinned on raw shared-memory bytes."""
    rng = np.random.RandomState(3)
    rows = 1000
    mat = torch.from_numpy(rng.uniform(-1, 1, (rows, 64)).astype(np.float32)).cuda().to(torch.bfloat16)
    wmat = torch.from_numpy(rng.uniform(-1, 1, (64, 64)).astype(np.float32)).cuda().to(torch.bfloat16)
    idx_np = rng.randint(0, rows, 128).astype(np.int32)
    idx_np[5] = rows - 1
    idx_np[6] = 0
    row0 = 256
    dump, d = ops.selftest_tma(mat, wmat, torch.from_numpy(idx_np).cuda(), row0)
    torch.cuda.synchronize()
    dump = dump.cpu().numpy()
    mat_b = mat.cpu().view(torch.int16).numpy().view(np.uint8).reshape(rows, 128)
    w_b = wmat.cpu().view(torch.int16).numpy().view(np.uint8).reshape(64, 128)

    def swizzled(rows_bytes):
        n = rows_bytes.shape[0]
        out = np.zeros((n, 8, 16), dtype=np.uint8)
        src = rows_bytes.reshape(n, 8, 16)
        for r in range(n):
            for j in range(8):
                out[r, j ^ (r & 7)] = src[r, j]
        return out.reshape(-1)

    assert np.array_equal(dump[:16384], swizzled(mat_b[row0:row0 + 128])), 'tile load layout'
    assert np.array_equal(dump[16384:32768], swizzled(mat_b[idx_np])), 'gather4 layout'
    assert np.array_equal(dump[32768:], swizzled(w_b)), 'weight tile layout'
    a = mat.float().cpu().numpy().astype(np.float64)
    ref = (a[row0:row0 + 128] + a[idx_np]) @ wmat.float().cpu().numpy().astype(np.float64).T
    assert np.max(np.abs(d.cpu().numpy() - ref)) < 1e-4 * np.max(np.abs(ref))
