"""GPU: DetectionMatching (A9) and the loss (A10) through the reference's op
signature vs the reference's own det_matching.cc build (oracle/_ref) and the
C++ restatement.  Bit-exact, including ties (same libstdc++ sort decisions)."""
import numpy as np
import pytest
import torch

from gossipnet_b200 import ops
from gossipnet_b200.nms_net.config import cfg
from gossipnet_b200.nms_net.matching_module import detection_matching
from oracle import det_matching_oracle as dm
from oracle import gnet_oracle as go
from tests.test_matching_oracle import random_case

pytestmark = pytest.mark.gpu
F32 = np.float32


def ref_match(iou, score, ignore):
    if dm.have_reference_build():
        return dm.ref_detection_matching(iou, score, ignore)
    return dm.detection_matching(iou, score, ignore)


def gpu_match(iou, score, ignore):
    return [t.cpu().numpy() for t in detection_matching(np.asarray(iou, F32),
                                                        np.asarray(score, F32),
                                                        np.asarray(ignore, bool))]


def test_known_answers(oracle_built):
    cases = [
        ([[0.9], [0.8]], [0.1, 0.7], [False]),
        ([[0.9], [0.8], [0.1]], [0.5, 0.6, 0.7], [True]),
        ([[0.9, 0.6]], [0.5], [True, False]),
        ([[0.6, 0.9]], [0.5], [True, True]),
        ([[0.7, 0.7]], [0.5], [False, False]),
        ([[0.49999], [0.5]], [0.9, 0.8], [False]),
    ]
    for iou, score, ignore in cases:
        got = gpu_match(iou, score, ignore)
        ref = ref_match(np.array(iou, F32), np.array(score, F32), np.array(ignore, bool))
        for a, b in zip(got, ref):
            assert np.array_equal(a, b)
    lab, w, a = gpu_match(np.zeros((3, 0), F32), [0.1, 0.2, 0.3], np.zeros(0, bool))
    assert lab.tolist() == [0, 0, 0] and w.tolist() == [1, 1, 1] and a.tolist() == [-1, -1, -1]
    assert lab.dtype == np.float32 and a.dtype == np.int32


@pytest.mark.parametrize('tie_scores,tie_iou', [(False, False), (True, False), (False, True),
                                               (True, True)])
def test_random_vs_reference_build(oracle_built, tie_scores, tie_iou):
    rs = np.random.RandomState(5)
    for n, g in [(1, 1), (17, 3), (40, 40), (300, 12), (1000, 40), (2500, 97), (10000, 100),
                 (33, 200)]:
        iou, score, ignore = random_case(rs, n, g, tie_scores, tie_iou)
        got = gpu_match(iou, score, ignore)
        ref = ref_match(iou, score, ignore)
        for a, b, name in zip(got, ref, ('labels', 'weights', 'assignment')):
            assert np.array_equal(a, b), (n, g, name)


def test_batched_images(oracle_built):
    rs = np.random.RandomState(9)
    shapes = [(300, 12), (0, 3), (1000, 40), (5, 0), (64, 64)]
    cases = [random_case(rs, n, g, True, True) for n, g in shapes]
    img_off = np.cumsum([0] + [n for n, _ in shapes]).astype(np.int32)
    gt_off = np.cumsum([0] + [g for _, g in shapes]).astype(np.int32)
    iou_off = np.cumsum([0] + [n * g for n, g in shapes]).astype(np.int64)
    cat = lambda i, dt: torch.from_numpy(np.concatenate([c[i].reshape(-1) for c in cases])
                                         .astype(dt)).cuda()
    lab, w, a = ops.detection_matching_batched(
        cat(0, F32), torch.from_numpy(iou_off).cuda(), cat(1, F32), cat(2, np.uint8),
        torch.from_numpy(img_off).cuda(), torch.from_numpy(gt_off).cuda(), 64)
    lab, w, a = lab.cpu().numpy(), w.cpu().numpy(), a.cpu().numpy()
    for i, (iou, score, ignore) in enumerate(cases):
        ref = ref_match(iou, score, ignore)
        s = slice(img_off[i], img_off[i + 1])
        assert np.array_equal(lab[s], ref[0]) and np.array_equal(w[s], ref[1])
        assert np.array_equal(a[s], ref[2])


def test_shape_errors_like_the_reference_op():
    with pytest.raises(ValueError):
        detection_matching(np.zeros(3, F32), np.zeros(3, F32), np.zeros(2, bool))
    with pytest.raises(ValueError):
        detection_matching(np.zeros((3, 2), F32), np.zeros(4, F32), np.zeros(2, bool))
    with pytest.raises(ValueError):
        detection_matching(np.zeros((3, 2), F32), np.zeros(3, F32), np.zeros(5, bool))


@pytest.mark.parametrize('normalize', [False, True])
def test_loss_vs_oracle(normalize):
    cfg.train.normalize_loss = normalize
    cfg.train.loss_multiplyer = 2.0
    rs = np.random.RandomState(1)
    n, g = 500, 20
    pred = rs.normal(0, 4, n).astype(F32)
    pred[:3] = [-30, 0, 30]
    labels = (rs.uniform(0, 1, n) < 0.2).astype(F32)
    weights = (rs.uniform(0, 1, n) < 0.9).astype(F32)
    assign = np.where(labels > 0, rs.randint(0, g, n), -1).astype(np.int32)
    gt_crowd = rs.uniform(0, 1, g) < 0.3
    gt_classes = rs.randint(1, 5, g).astype(np.int32)
    cw = rs.uniform(0.5, 2, 5).astype(F32)
    ref = go.loss(pred, labels, weights, assign, gt_crowd, gt_classes, cw, cfg)
    d = lambda a, dt=None: torch.from_numpy(np.ascontiguousarray(a if dt is None else a.astype(dt))).cuda()
    w_io = d(weights)
    loss_out, dlogit = ops.loss_fwd(d(pred), d(labels), w_io, d(assign), d(gt_crowd, np.uint8),
                                    d(gt_classes), d(np.array([0, n], np.int32)),
                                    d(np.array([0, g], np.int32)), d(cw), normalize, 2.0,
                                    want_grad=True)
    lo = loss_out.cpu().numpy()[0]
    assert np.array_equal(w_io.cpu().numpy(), ref['weights'])
    assert abs(lo[0] - ref['loss_unnormed']) <= 1e-5 * abs(ref['loss_unnormed'])
    assert abs(lo[1] - ref['loss_normed']) <= 1e-5 * abs(ref['loss_normed'])
    assert abs(lo[2] - ref['loss']) <= 1e-5 * abs(ref['loss'])
    # analytic gradient of the reported loss w.r.t. the logits
    x = pred.astype(np.float64)
    sig = 1 / (1 + np.exp(-x))
    scale = 2.0 * (1.0 / n if normalize else 1.0)
    assert np.allclose(dlogit.cpu().numpy(), scale * ref['weights'] * (sig - labels), atol=1e-6)
