"""GPU: the tensor-core fully connected layer and its weight gradient (gn_fc_tc.cu, the
training step's GEMMs) against float64 products: y = act(res + (x.mask) W + b), the input
gradient through the transposed operand image, dW += x^T (dy.mask), db += colsum.
bf16x3 keeps fp32 semantics: 1e-5 relative (max-normalised) like the fused kernels."""
import numpy as np
import pytest
import torch

from gossipnet_b200 import ops

pytestmark = pytest.mark.gpu
TOL = 2e-5


def images(w, transposed):
    """Operand image of one weight matrix through gn_prepare_fc_images."""
    k, n = w.shape
    kpad = ((n if transposed else k) + 15) // 16 * 16
    N = k if transposed else n
    nbytes = 2 * (kpad // 8) * N * 16
    table = torch.tensor([[0, k, n, 0, kpad, 1 if transposed else 0]], dtype=torch.int32).cuda()
    img = torch.zeros(nbytes, dtype=torch.uint8, device='cuda')
    ops.prepare_fc_images(w.reshape(-1).contiguous(), table, img)
    return img


def rel(a, b):
    return float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), 1e-12))


@pytest.mark.parametrize('rows,k,n', [(1000, 9, 256), (777, 256, 256), (5000, 256, 32),
                                      (300, 96, 64), (129, 64, 128), (1, 128, 32), (40000, 64, 64)])
@pytest.mark.parametrize('relu,res', [(True, False), (False, False), (True, True)])
def test_fc_fwd_tc(rows, k, n, relu, res):
    rs = np.random.RandomState(rows + k + n)
    x = rs.normal(0, 1, (rows, k)).astype(np.float32)
    w = rs.normal(0, 0.2, (k, n)).astype(np.float32)
    b = rs.normal(0, 0.1, n).astype(np.float32)
    r = rs.normal(0, 1, (rows, n)).astype(np.float32) if res else None
    d = lambda a: None if a is None else torch.from_numpy(a).cuda()
    rows_dev = torch.tensor([rows - rows // 7], dtype=torch.int32).cuda()
    live = rows - rows // 7
    out = torch.full((rows, n), -7.0, device='cuda')
    ops.fc_fwd_tc(d(x), images(d(w), False), k, n, d(b), relu, residual=d(r), out=out,
                  rows_dev=rows_dev)
    ref = x.astype(np.float64) @ w.astype(np.float64) + b
    if res:
        ref = ref + r
    if relu:
        ref = np.maximum(ref, 0)
    got = out.cpu().numpy()
    assert rel(got[:live], ref[:live]) < TOL
    assert np.all(got[live:] == -7.0)          # rows past the device-side count untouched


@pytest.mark.parametrize('rows,k,n', [(2000, 256, 256), (513, 96, 64), (900, 128, 32), (64, 32, 128)])
def test_fc_input_gradient_with_fused_relu_mask(rows, k, n):
    """dx = (dy . (y > 0)) @ W^T through the transposed image and the fused mask."""
    rs = np.random.RandomState(rows + n)
    dy = rs.normal(0, 1, (rows, n)).astype(np.float32)
    y = rs.normal(0, 1, (rows, n)).astype(np.float32)
    w = rs.normal(0, 0.2, (k, n)).astype(np.float32)
    d = lambda a: torch.from_numpy(a).cuda()
    got = ops.fc_fwd_tc(d(dy), images(d(w), True), n, k, None, False, mask=d(y)).cpu().numpy()
    ref = (dy * (y > 0)).astype(np.float64) @ w.astype(np.float64).T
    assert rel(got, ref) < TOL


@pytest.mark.parametrize('rows,k,n', [(1000, 9, 256), (3000, 256, 256), (70000, 256, 32),
                                      (515, 96, 64), (129, 64, 128), (31, 128, 32), (8000, 64, 64)])
@pytest.mark.parametrize('masked', [False, True])
def test_fc_bwd_weight_tc(rows, k, n, masked):
    rs = np.random.RandomState(rows + k)
    x = rs.normal(0, 1, (rows, k)).astype(np.float32)
    dy = rs.normal(0, 1, (rows, n)).astype(np.float32)
    y = rs.normal(0, 1, (rows, n)).astype(np.float32)
    dw0 = rs.normal(0, 1, (k, n)).astype(np.float32)
    db0 = rs.normal(0, 1, n).astype(np.float32)
    d = lambda a: torch.from_numpy(a).cuda()
    live = rows - rows // 5
    dw, db = d(dw0), d(db0)
    ops.fc_bwd_weight_tc(d(x), d(dy), dw, db, rows_dev=torch.tensor([live], dtype=torch.int32).cuda(),
                         mask=d(y) if masked else None)
    g = dy * (y > 0) if masked else dy
    ref_w = dw0 + x[:live].astype(np.float64).T @ g[:live].astype(np.float64)
    ref_b = db0 + g[:live].astype(np.float64).sum(0)
    assert rel(dw.cpu().numpy(), ref_w) < TOL
    assert rel(db.cpu().numpy(), ref_b) < TOL
