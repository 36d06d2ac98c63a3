"""CPU: the C restatement of ROI max pooling against the reference's OWN
roi_pooling_op.cc (compiled unmodified into oracle/_ref) - bit-exact, forward
and backward - plus hand-derived known answers."""
import numpy as np
import pytest

from oracle import roi_pool_oracle as rp


def random_case(rs, b, h, w, c, r, scale=1.0 / 16):
    data = rs.normal(0, 1, (b, h, w, c)).astype(np.float32)
    x1 = rs.uniform(-20, w / scale, r)
    y1 = rs.uniform(-20, h / scale, r)
    x2 = x1 + rs.uniform(-10, 0.7 * w / scale, r)     # some malformed (x2 < x1) on purpose
    y2 = y1 + rs.uniform(-10, 0.7 * h / scale, r)
    rois = np.stack([rs.randint(0, b, r).astype(np.float64), x1, y1, x2, y2], 1).astype(np.float32)
    return data, rois


def test_known_answers():
    data = np.arange(1 * 4 * 4 * 1, dtype=np.float32).reshape(1, 4, 4, 1)
    top, arg = rp.roi_pool(data, np.array([[0, 0, 0, 3, 3]], np.float32), 2, 2, 1.0)
    assert top.reshape(2, 2).tolist() == [[5, 7], [13, 15]]
    assert arg.reshape(2, 2).tolist() == [[5, 7], [13, 15]]
    # roi entirely outside the map: empty bins -> 0, argmax -1
    top, arg = rp.roi_pool(data, np.array([[0, 10, 10, 12, 12]], np.float32), 2, 2, 1.0)
    assert np.all(top == 0) and np.all(arg == -1)
    # gradient goes to the argmax cell only
    g = rp.roi_pool_grad(data, np.array([[0, 0, 0, 3, 3]], np.float32),
                         np.array([5, 7, 13, 15], np.int32).reshape(1, 2, 2, 1),
                         np.array([1, 2, 3, 4], np.float32).reshape(1, 2, 2, 1), 2, 2, 1.0)
    want = np.zeros(16, np.float32)
    want[[5, 7, 13, 15]] = [1, 2, 3, 4]
    assert g.reshape(-1).tolist() == want.tolist()


@pytest.mark.parametrize('shape', [(1, 8, 9, 5, 7), (2, 16, 20, 32, 40), (1, 38, 63, 16, 100)])
def test_restatement_equals_reference_build(shape):
    if not rp.have_reference_build():
        pytest.skip('oracle/_ref not built (no /root/reference here)')
    rs = np.random.RandomState(sum(shape))
    b, h, w, c, r = shape
    data, rois = random_case(rs, b, h, w, c, r)
    for ph, pw in [(7, 7), (2, 3)]:
        top, arg = rp.roi_pool(data, rois, ph, pw, 1.0 / 16)
        rtop, rarg = rp.ref_roi_pool(data, rois, ph, pw, 1.0 / 16)
        assert np.array_equal(arg, rarg) and np.array_equal(top.view(np.uint32), rtop.view(np.uint32))
        grad = rs.normal(0, 1, top.shape).astype(np.float32)
        g = rp.roi_pool_grad(data, rois, arg, grad, ph, pw, 1.0 / 16)
        rg = rp.ref_roi_pool_grad(data, rois, rarg, grad, ph, pw, 1.0 / 16)
        assert np.array_equal(g.view(np.uint32), rg.view(np.uint32))


def test_reference_build_rejects_bad_rank():
    if not rp.have_reference_build():
        pytest.skip('oracle/_ref not built')
    with pytest.raises(ValueError):
        rp.ref_roi_pool(np.zeros((1, 4, 4, 1), np.float32), np.zeros(5, np.float32), 2, 2, 1.0)
