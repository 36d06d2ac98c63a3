"""GPU: IoU (A1+A2) and neighbor build (A3) through the C ABI vs the oracle.
Bit-exact: these feed the threshold and the matching."""
import numpy as np
import pytest
import torch

from gossipnet_b200 import ops, synthetic
from oracle import gnet_oracle as go

pytestmark = pytest.mark.gpu
F32 = np.float32


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def bits(a):
    return np.ascontiguousarray(a, dtype=F32).view(np.uint32)


@pytest.mark.parametrize('n', [1, 3, 37, 300, 1000, 1003])
def test_det_det_iou_bit_exact(n):
    d = synthetic.make_image(n, 1, image_index=n)['dets']
    ref = go.iou(go.xyxy_to_boxdata(d), go.xyxy_to_boxdata(d))
    got = ops.iou_dense(dev(d), dev(d)).cpu().numpy()
    assert np.array_equal(bits(got), bits(ref))
    assert np.all(np.diag(got) == 1.0)
    assert np.array_equal(got, got.T)


@pytest.mark.parametrize('n', [128, 129, 1000, 1003, 4000, 10000])
def test_symmetric_iou_kernel_bit_exact(n):
    """The SAME tensor on both sides selects iou_symmetric_kernel (gn_iou.cu: upper-triangle
    tiles computed once, stored twice) - the kernel `roofline_iou` is quoted on.  Bit-exact
    against the oracle (network.py:474-511) up to the stress size."""
    t = dev(synthetic.make_image(n, 1, image_index=n)['dets'])
    got = ops.iou_dense(t, t).cpu().numpy()
    d = t.cpu().numpy()
    ref = go.iou(go.xyxy_to_boxdata(d), go.xyxy_to_boxdata(d))
    assert np.array_equal(bits(got), bits(ref))
    del ref
    assert np.all(np.diag(got) == 1.0)
    # and the general kernel (two buffers) gives the same bits
    assert np.array_equal(bits(ops.iou_dense(t, t.clone()).cpu().numpy()), bits(got))


@pytest.mark.parametrize('n', [130, 512, 1500])
def test_symmetric_iou_dense_overlaps(n):
    """Heavily overlapping boxes (every pair intersects): the deferred-division queue of
    iou_symmetric_kernel overflows and the in-place path takes over; still bit-exact.  Also
    degenerate boxes (zero area -> 0/0 = NaN like the reference arithmetic)."""
    rs = np.random.RandomState(n)
    c = rs.uniform(400, 600, (n, 2))
    wh = rs.uniform(150, 300, (n, 2))
    d = np.concatenate([c - wh / 2, c + wh / 2], axis=1).astype(F32)
    d[3] = [10, 10, 10, 20]          # zero-area boxes
    d[7] = [10, 10, 10, 20]
    t = dev(d)
    got = ops.iou_dense(t, t).cpu().numpy()
    with np.errstate(invalid='ignore', divide='ignore'):
        ref = go.iou(go.xyxy_to_boxdata(d), go.xyxy_to_boxdata(d))
    nan = np.isnan(ref)      # 0/0 of the zero-area boxes: NaN on both sides (payload bits differ
    assert nan.sum() == 4 and np.array_equal(np.isnan(got), nan)      # between x86 and the GPU)
    assert np.array_equal(bits(got)[~nan], bits(ref)[~nan])
    # half dense / half sparse: queue partly filled, some threads in place
    d2 = d.copy()
    d2[n // 2:, :2] += 5000
    d2[n // 2:, 2:] += 5000
    d2[n // 2:] += (np.arange(n - n // 2)[:, None] * 400).astype(F32)
    t2 = dev(d2)
    got2 = ops.iou_dense(t2, t2).cpu().numpy()
    with np.errstate(invalid='ignore', divide='ignore'):
        ref2 = go.iou(go.xyxy_to_boxdata(d2), go.xyxy_to_boxdata(d2))
    assert np.array_equal(bits(got2)[~np.isnan(ref2)], bits(ref2)[~np.isnan(ref2)])


def test_iou_known_answers():
    d = np.array([[0, 0, 10, 10], [5, 0, 15, 10], [100, 100, 110, 120]], F32)
    got = ops.iou_dense(dev(d), dev(d)).cpu().numpy()
    assert got[0, 1] == F32(50.0) / F32(150.0)
    assert got[0, 2] == 0.0 and not np.signbit(got[0, 2])


@pytest.mark.parametrize('n,g,multi', [(300, 12, False), (257, 7, True), (1000, 40, True),
                                       (50, 1, False), (5, 0, False)])
def test_det_gt_iou_crowd_and_class_mask(n, g, multi):
    rs = np.random.RandomState(n + g)
    C = 80 if multi else 1
    img = synthetic.make_image(n, C, image_index=g)
    d = img['dets']
    gt = synthetic.make_image(max(g, 1), C, image_index=100 + g)['dets'][:g]
    crowd = rs.uniform(0, 1, g) < 0.4
    dcls = rs.randint(1, 4, n).astype(np.int32)
    gcls = rs.randint(1, 4, g).astype(np.int32)
    ref = go.iou(go.xyxy_to_boxdata(d), go.xyxy_to_boxdata(gt), crowd)
    if multi:
        ref = go.class_mask_iou(ref, dcls, gcls)
    got = ops.iou_dense(dev(d), dev(gt.reshape(-1, 4)), crowd=dev(crowd.astype(np.uint8)),
                        a_cls=dev(dcls) if multi else None,
                        b_cls=dev(gcls) if multi else None).cpu().numpy()
    assert got.shape == (n, g)
    assert np.array_equal(bits(got), bits(ref.reshape(n, g)))


def test_batched_iou_matches_per_image():
    imgs = [synthetic.make_image(200, 1, image_index=i)['dets'] for i in range(5)]
    a = dev(np.stack(imgs))
    got = ops.iou_dense(a, a).cpu().numpy()
    for i, d in enumerate(imgs):
        ref = go.iou(go.xyxy_to_boxdata(d), go.xyxy_to_boxdata(d))
        assert np.array_equal(bits(got[i]), bits(ref))


def build_pairs(dets_list, thresh=0.2, masks=False):
    n = [d.shape[0] for d in dets_list]
    off = np.zeros(len(n) + 1, np.int32)
    np.cumsum(n, out=off[1:])
    dets = dev(np.concatenate(dets_list).reshape(-1, 4))
    img_off = dev(off)
    if masks:      # the shipped variant: division-free count pass + per-row hit masks
        T = int(off[-1])
        stride = max(1, (max(n) + 31) // 32)
        degree = torch.empty(T, dtype=torch.int32, device='cuda')
        hit = torch.full((T, stride), -1, dtype=torch.int32, device='cuda')
        ops.neighbor_count_masks(dets, img_off, thresh, stride, degree, hit)
        row_ptr = ops.exclusive_scan(degree)
        P = int(row_ptr[-1].item())
        pc = torch.empty(P, dtype=torch.int32, device='cuda')
        pn = torch.empty(P, dtype=torch.int32, device='cuda')
        pi = torch.empty(P, dtype=torch.float32, device='cuda')
        ops.neighbor_fill_masks(dets, img_off, row_ptr, P, hit, stride, pc, pn, pi)
        return off, row_ptr.cpu().numpy(), pc.cpu().numpy(), pn.cpu().numpy(), pi.cpu().numpy()
    degree = ops.neighbor_count(dets, img_off, thresh)
    row_ptr = ops.exclusive_scan(degree)
    P = int(row_ptr[-1].item())
    pc = torch.empty(P, dtype=torch.int32, device='cuda')
    pn = torch.empty(P, dtype=torch.int32, device='cuda')
    pi = torch.empty(P, dtype=torch.float32, device='cuda')
    ovf = torch.zeros(1, dtype=torch.int32, device='cuda')
    ops.neighbor_fill(dets, img_off, thresh, row_ptr, P, pc, pn, pi, ovf)
    assert int(ovf.item()) == 0
    return off, row_ptr.cpu().numpy(), pc.cpu().numpy(), pn.cpu().numpy(), pi.cpu().numpy()


@pytest.mark.parametrize('masks', [False, True])
@pytest.mark.parametrize('sizes', [[1], [40], [300], [1000], [17, 1, 255, 64, 3], [1000] * 4,
                                   [2000], [0, 5, 0, 7], [3, 2, 1, 1, 1, 6, 33, 2]])
def test_neighbor_lists_bit_exact(sizes, masks):
    dets_list = [synthetic.make_image(max(n, 1), 1, image_index=i)['dets'][:n]
                 for i, n in enumerate(sizes)]
    off, row_ptr, pc, pn, pi = build_pairs(dets_list, masks=masks)
    want_c, want_n, want_iou = [], [], []
    for i, d in enumerate(dets_list):
        bd = go.xyxy_to_boxdata(d)
        m = go.iou(bd, bd)
        pairs = go.neighbor_pairs(m, 0.2)
        want_c.append(pairs[:, 0] + off[i])
        want_n.append(pairs[:, 1] + off[i])
        want_iou.append(m[pairs[:, 0], pairs[:, 1]])
    want_c = np.concatenate(want_c) if want_c else np.zeros(0, np.int64)
    want_n = np.concatenate(want_n)
    assert np.array_equal(pc.astype(np.int64), want_c)
    assert np.array_equal(pn.astype(np.int64), want_n)
    assert np.array_equal(bits(pi), bits(np.concatenate(want_iou)))
    # CSR consistency
    T = int(off[-1])
    assert row_ptr.shape == (T + 1,) and row_ptr[0] == 0 and row_ptr[-1] == pc.shape[0]
    assert np.array_equal(np.bincount(pc, minlength=T), np.diff(row_ptr))


def test_threshold_edge_inclusive_on_gpu():
    # find a pair of boxes whose IoU is exactly float32(0.2): inter 2x5=10 ... search
    thr = F32(0.2)
    found = None
    for w in range(2, 60):
        a = np.array([[0, 0, 10, 10], [10 - w * 0.25, 0, 20 - w * 0.25, 10]], F32)
        m = go.iou(go.xyxy_to_boxdata(a), go.xyxy_to_boxdata(a))
        if m[0, 1] == thr:
            found = a
            break
    if found is None:
        # inter/union = 0.2 exactly: inter 20, union 100 -> boxes 10x6 overlapping 10/3? use ints
        found = np.array([[0, 0, 12, 10], [8, 0, 20, 10]], F32)  # inter 40, union 200
    m = go.iou(go.xyxy_to_boxdata(found), go.xyxy_to_boxdata(found))
    assert m[0, 1] == thr
    for masks in (False, True):     # the division-free test must agree on the exact edge
        _, _, pc, pn, _ = build_pairs([found], masks=masks)
        assert list(zip(pc.tolist(), pn.tolist())) == [(0, 0), (0, 1), (1, 0), (1, 1)]
        _, _, pc, pn, _ = build_pairs([found], thresh=float(np.nextafter(thr, F32(1))), masks=masks)
        assert list(zip(pc.tolist(), pn.tolist())) == [(0, 0), (1, 1)]


def test_division_free_threshold_agrees_near_the_edge():
    """Pairs engineered to land within a few ulps of the threshold on both sides: the mask
    variant (fast accept / reject + exact fallback) equals thresholding the dense IoU."""
    rs = np.random.RandomState(3)
    boxes = []
    for _ in range(400):
        w, h = rs.uniform(20, 200, 2)
        # two boxes of equal size shifted along x so that iou = (w - s) / (w + s) ~ 0.2
        s = w * (1 - 0.2) / (1 + 0.2) * (1 + rs.uniform(-3e-7, 3e-7))
        x, y = rs.uniform(0, 500, 2)
        boxes += [[x, y, x + w, y + h], [x + s, y, x + s + w, y + h]]
    d = np.array(boxes, F32)
    for thresh in (0.2, float(np.nextafter(F32(0.2), F32(1))), float(np.nextafter(F32(0.2), F32(0)))):
        off, _, pc, pn, pi = build_pairs([d], thresh=thresh, masks=True)
        m = go.iou(go.xyxy_to_boxdata(d), go.xyxy_to_boxdata(d))
        want = np.argwhere(m >= F32(thresh))
        assert np.array_equal(np.stack([pc, pn], axis=1), want)
        assert np.array_equal(bits(pi), bits(m[want[:, 0], want[:, 1]]))


def test_fill_reports_overflow_and_stays_in_bounds():
    d = synthetic.make_image(300, 1)['dets']
    dets, img_off = dev(d), dev(np.array([0, 300], np.int32))
    row_ptr = ops.exclusive_scan(ops.neighbor_count(dets, img_off, 0.2))
    P = int(row_ptr[-1].item())
    cap = P // 2
    guard = 64
    pc = torch.full((cap + guard,), -7, dtype=torch.int32, device='cuda')
    pn = torch.full((cap + guard,), -7, dtype=torch.int32, device='cuda')
    pi = torch.full((cap + guard,), -7.0, dtype=torch.float32, device='cuda')
    ovf = torch.zeros(1, dtype=torch.int32, device='cuda')
    ops.neighbor_fill(dets, img_off, 0.2, row_ptr, cap, pc, pn, pi, ovf)
    assert int(ovf.item()) == 1
    assert torch.all(pc[cap:] == -7) and torch.all(pn[cap:] == -7) and torch.all(pi[cap:] == -7)


def test_argument_validation():
    a = torch.zeros((4, 4), dtype=torch.float32)
    with pytest.raises(ValueError):
        ops.iou_dense(a, a)  # CPU tensor: no CPU fallback
    with pytest.raises(ValueError):
        ops.iou_dense(a.cuda().double(), a.cuda().double())


@pytest.mark.parametrize('n', [1, 5, 4095, 4096, 16383, 16384, 16385, 64000, 200001])
@pytest.mark.parametrize('shift', [0, 1])
def test_exclusive_scan(n, shift):
    """gn_exclusive_scan (row_ptr of the neighbor lists, network.py:192-195's ordering) against
    numpy, over round / segment boundaries and with buffers that are not 16-byte aligned."""
    rs = np.random.RandomState(n)
    deg = rs.randint(0, 90, size=n).astype(np.int32)
    src = torch.zeros(n + shift, dtype=torch.int32, device='cuda')
    src[shift:] = torch.from_numpy(deg).cuda()
    dst = torch.full((n + 1 + shift,), -7, dtype=torch.int32, device='cuda')
    got = ops.exclusive_scan(src[shift:], out=dst[shift:]).cpu().numpy()
    ref = np.concatenate([[0], np.cumsum(deg.astype(np.int64))]).astype(np.int32)
    assert np.array_equal(got, ref)
    if shift:
        assert int(dst[0]) == -7
