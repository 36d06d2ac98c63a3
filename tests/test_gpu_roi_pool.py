"""GPU: RoiPool / RoiPoolGrad through the reference's op signatures vs the
reference's own CPU kernels (oracle/_ref) / the C restatement.  Bit-exact."""
import numpy as np
import pytest
import torch

from gossipnet_b200.nms_net.roi_pooling_layer import roi_pooling_op, roi_pooling_op_grad
from oracle import roi_pool_oracle as rp
from tests.test_roi_pool_oracle import random_case

pytestmark = pytest.mark.gpu


def cpu_fwd(*a):
    return rp.ref_roi_pool(*a) if rp.have_reference_build() else rp.roi_pool(*a)


def cpu_bwd(*a):
    return rp.ref_roi_pool_grad(*a) if rp.have_reference_build() else rp.roi_pool_grad(*a)


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize('shape', [(1, 8, 9, 5, 7), (2, 16, 20, 32, 40), (1, 38, 63, 64, 300),
                                   (3, 12, 12, 1024, 25), (1, 5, 5, 3, 600)])
@pytest.mark.parametrize('pool', [(7, 7), (2, 3)])
def test_forward_backward_bit_exact(shape, pool):
    rs = np.random.RandomState(sum(shape) + pool[0])
    b, h, w, c, r = shape
    data, rois = random_case(rs, b, h, w, c, r)
    ph, pw = pool
    top, arg = roi_pooling_op.roi_pool(dev(data), dev(rois), ph, pw, 1.0 / 16)
    rtop, rarg = cpu_fwd(data, rois, ph, pw, 1.0 / 16)
    assert top.dtype == torch.float32 and arg.dtype == torch.int32
    assert np.array_equal(arg.cpu().numpy(), rarg)
    assert np.array_equal(top.cpu().numpy().view(np.uint32), rtop.view(np.uint32))
    grad = rs.normal(0, 1, rtop.shape).astype(np.float32)
    g = roi_pooling_op.roi_pool_grad(dev(data), dev(rois), arg, dev(grad), ph, pw, 1.0 / 16)
    rg = cpu_bwd(data, rois, rarg, grad, ph, pw, 1.0 / 16)
    assert np.array_equal(g.cpu().numpy().view(np.uint32), rg.view(np.uint32))


def test_full_size_feature_map_round_trip_properties():
    """conv5-sized map (38 x 63 x 1024, stride 16) with N=1000 rois, 7x7 bins: the
    gradient of sum(top) puts exactly one unit per non-empty pooled cell."""
    rs = np.random.RandomState(1)
    data, _ = random_case(rs, 1, 38, 63, 1024, 1)
    # well-formed rois (the reference drops the gradient of malformed ones: its
    # backward tests start <= w <= end, which a forced-1x1 roi never passes)
    x1 = rs.uniform(0, 900, 1000)
    y1 = rs.uniform(0, 500, 1000)
    rois = np.stack([np.zeros(1000), x1, y1, x1 + rs.uniform(16, 400, 1000),
                     y1 + rs.uniform(16, 300, 1000)], 1).astype(np.float32)
    top, arg = roi_pooling_op.roi_pool(dev(data), dev(rois), 7, 7, 1.0 / 16)
    ones = torch.ones_like(top)
    g = roi_pooling_op.roi_pool_grad(dev(data), dev(rois), arg, ones, 7, 7, 1.0 / 16)
    assert float(g.double().sum()) == float((arg >= 0).sum())
    flat = dev(data).reshape(-1)
    sel = arg.reshape(-1) >= 0
    assert torch.equal(top.reshape(-1)[sel], flat[arg.reshape(-1)[sel].long()])
    assert torch.all(top.reshape(-1)[~sel] == 0)


def test_autograd_wiring():
    rs = np.random.RandomState(2)
    data, rois = random_case(rs, 1, 10, 10, 8, 6)
    x = dev(data).requires_grad_(True)
    top, arg = roi_pooling_op_grad.roi_pool_with_grad(x, dev(rois), 3, 3, 1.0 / 16)
    top.sum().backward()
    want = cpu_bwd(data, rois, arg.cpu().numpy(), np.ones(top.shape, np.float32), 3, 3, 1.0 / 16)
    assert np.array_equal(x.grad.cpu().numpy(), want)


def test_argument_errors_like_the_reference_op():
    with pytest.raises(ValueError):
        roi_pooling_op.roi_pool(torch.zeros((4, 4, 1), device='cuda'), torch.zeros((1, 5), device='cuda'), 2, 2, 1.0)
    with pytest.raises(ValueError):
        roi_pooling_op.roi_pool(torch.zeros((1, 4, 4, 1), device='cuda'), torch.zeros(5, device='cuda'), 2, 2, 1.0)
    with pytest.raises(ValueError):
        roi_pooling_op.roi_pool(torch.zeros((1, 4, 4, 1), device='cuda'), torch.zeros((1, 5), device='cuda'), -1, 2, 1.0)
