"""CPU: hand-derived known-answer tests for the oracle (SURVEY.md §8c list)."""
import numpy as np

from gossipnet_b200.nms_net.config import cfg
from oracle import det_matching_oracle as dm
from oracle import gnet_oracle as go

F32 = np.float32


def boxes(*rows):
    return go.xyxy_to_boxdata(np.array(rows, dtype=F32))


def test_iou_known_values():
    a = boxes([0, 0, 10, 10], [5, 0, 15, 10], [100, 100, 110, 120])
    m = go.iou(a, a)
    assert m.dtype == F32
    assert m[0, 1] == F32(50.0) / F32(150.0)          # inter 50, union 150
    assert np.all(np.diag(m) == F32(1.0))             # self IoU exactly 1
    assert np.array_equal(m, m.T)                     # bitwise symmetric
    assert m[0, 2] == 0.0 and not np.signbit(m[0, 2])  # disjoint -> +0.0


def test_iou_crowd_column_is_intersection_over_det_area():
    a = boxes([0, 0, 10, 10])
    b = boxes([5, 0, 25, 10], [5, 0, 25, 10])
    m = go.iou(a, b, crowd=np.array([False, True]))
    assert m[0, 0] == F32(50.0) / F32((100.0 + 200.0) - 50.0)
    assert m[0, 1] == F32(50.0) / F32(100.0)


def test_threshold_edge_is_inclusive():
    # engineered pair with iou exactly float32(0.2): inter 20 / union 100
    a = boxes([0, 0, 10, 6], [8, 0, 18, 6])   # inter 2*6=12, union 108 -> no
    thr = F32(0.2)
    m = np.array([[1.0, thr], [np.nextafter(thr, F32(0)), 1.0]], dtype=F32)
    pairs = go.neighbor_pairs(m, 0.2)
    assert pairs.tolist() == [[0, 0], [0, 1], [1, 1]]
    assert go.neighbor_pairs(go.iou(a, a), 0.2).tolist() == [[0, 0], [1, 1]]


def test_pair_order_is_row_major_with_diagonal():
    rs = np.random.RandomState(0)
    m = rs.uniform(0, 1, (17, 17)).astype(F32)
    np.fill_diagonal(m, 1.0)
    pairs = go.neighbor_pairs(m, 0.2)
    assert pairs.dtype == np.int64
    assert np.array_equal(pairs, np.argwhere(m >= F32(0.2)))
    key = pairs[:, 0] * 17 + pairs[:, 1]
    assert np.all(np.diff(key) > 0)
    assert all([i, i] in pairs.tolist() for i in range(17))


def test_geometry_identical_and_doubled_width():
    d = np.array([[10, 10, 30, 50], [10, 10, 30, 50], [10, 10, 50, 50]], dtype=F32)
    bd = go.xyxy_to_boxdata(d)
    m = go.iou(bd, bd)
    s = np.array([0.25, 0.25, 0.5], dtype=F32)
    pairs = np.array([[0, 1], [0, 2]], dtype=np.int64)
    f = go.geometry_feats(bd, m, s, np.ones(3, np.int32), pairs, 1)
    assert f.shape == (2, 9)
    assert np.array_equal(f[0], np.array([0.25, 0.25, 1, 0, 0, 0, 0, 0, 0], dtype=F32))
    # neighbour twice as wide, same height: w_diff = 1, h_diff = 0, aspect_diff = 1
    assert abs(f[1, 6] - 1.0) < 1e-6 and f[1, 7] == 0.0 and abs(f[1, 8] - 1.0) < 1e-6
    # x_dist = (30 - 20) / ((20 + 40) / 2)
    assert abs(f[1, 3] - 10.0 / 30.0) < 1e-7 and f[1, 4] == 0.0


def test_geometry_multiclass_one_hot_columns():
    d = np.array([[0, 0, 10, 10], [1, 1, 11, 11]], dtype=F32)
    bd = go.xyxy_to_boxdata(d)
    m = go.iou(bd, bd)
    f = go.geometry_feats(bd, m, np.array([0.3, 0.7], F32), np.array([2, 3], np.int32),
                          np.array([[0, 1]], np.int64), 4)
    assert f.shape == (1, 2 * 4 + 7)
    assert f[0, :4].tolist() == [0, F32(0.3), 0, 0]      # class 2 -> column 1
    assert f[0, 4:8].tolist() == [0, 0, F32(0.7), 0]     # class 3 -> column 2


def test_block_tiny_graph_with_isolated_detection():
    """3 detections, det 2 isolated (only its self pair): its n_feats row is
    zeroed; weights chosen so every stage is hand-computable."""
    cfg.gnet.shortcut_dim = 2
    cfg.gnet.reduced_dim = 1
    cfg.gnet.pairfeat_dim = 1
    cfg.gnet.num_block_pw_fc = 1
    cfg.gnet.num_block_fc = 1
    p = {
        'gnet/block1/reduce_dim/weights': np.array([[1.0], [0.0]], F32),
        'gnet/block1/reduce_dim/biases': np.zeros(1, F32),
        # x = [pw, c_feat, n_feat]; out = pw + 10*c + 100*n
        'gnet/block1/pw_fc1/weights': np.array([[1.0], [10.0], [100.0]], F32),
        'gnet/block1/pw_fc1/biases': np.zeros(1, F32),
        'gnet/block1/fc1/weights': np.array([[1.0, -1.0]], F32),
        'gnet/block1/fc1/biases': np.zeros(2, F32),
    }
    infeats = np.array([[1, 5], [2, 5], [3, 5]], F32)
    pair_c = np.array([0, 0, 1, 1, 2])
    pair_n = np.array([0, 1, 0, 1, 2])
    pw = np.array([[0.5], [0.25], [0.125], [0.0625], [0.03125]], F32)
    out = go.block(1, infeats, pair_c, pair_n, pw, p, cfg)
    # pair rows: (0,0): .5+10 ; (0,1): .25+10+200 ; (1,0): .125+20+100 ; (1,1): .0625+20 ; (2,2): .03125+30
    pooled = np.array([210.25, 120.125, 30.03125], F32)
    want = np.maximum(infeats + np.stack([pooled, -pooled], 1), 0)
    assert np.array_equal(out, want)


def _match(iou, score, ignore):
    return dm.detection_matching(np.array(iou, F32), np.array(score, F32),
                                 np.array(ignore, bool))


def test_matching_known_answers(oracle_built):
    # (i) two dets on one GT: the higher score gets it
    lab, w, a = _match([[0.9], [0.8]], [0.1, 0.7], [False])
    assert lab.tolist() == [0, 1] and w.tolist() == [1, 1] and a.tolist() == [-1, 0]
    # (ii) a crowd GT matches many dets, label 1 weight 0
    lab, w, a = _match([[0.9], [0.8], [0.1]], [0.5, 0.6, 0.7], [True])
    assert lab.tolist() == [1, 1, 0] and w.tolist() == [0, 0, 1] and a.tolist() == [0, 0, -1]
    # (iii) regular GT (0.6) beats a crowd GT (0.9): scan breaks at the first crowd
    lab, w, a = _match([[0.9, 0.6]], [0.5], [True, False])
    assert lab.tolist() == [1] and w.tolist() == [1] and a.tolist() == [1]
    # (iv) two crowd GTs >= 0.5: the first in GT order wins, not the best
    lab, w, a = _match([[0.6, 0.9]], [0.5], [True, True])
    assert a.tolist() == [0] and w.tolist() == [0]
    # (v) equal IoU with two regular GTs: the later one wins (>=)
    lab, w, a = _match([[0.7, 0.7]], [0.5], [False, False])
    assert a.tolist() == [1]
    # (vi) no GT at all
    lab, w, a = _match(np.zeros((3, 0)), [0.1, 0.2, 0.3], [])
    assert lab.tolist() == [0, 0, 0] and w.tolist() == [1, 1, 1] and a.tolist() == [-1, -1, -1]
    # below the hard-coded 0.5 threshold: never matched; exactly 0.5 is
    lab, w, a = _match([[0.49999], [0.5]], [0.9, 0.8], [False])
    assert a.tolist() == [-1, 0]


def test_sigmoid_ce_closed_form():
    x = np.array([-30.0, 0.0, 30.0], F32)
    for z in (0.0, 1.0):
        got = go.sigmoid_ce(x, np.full(3, z, F32)).astype(np.float64)
        want = np.log1p(np.exp(-np.abs(x.astype(np.float64)))) + np.maximum(x, 0) - x * z
        assert np.allclose(got, want, rtol=1e-6, atol=1e-12)
    assert go.sigmoid_ce(np.array([0.0], F32), np.array([1.0], F32))[0] == F32(np.log(2.0))


def test_loss_class_weighting_and_crowd():
    cfg.train.normalize_loss = False
    pred = np.array([0.0, 0.0, 0.0, 0.0], F32)
    labels = np.array([1, 1, 0, 1], F32)
    weights = np.array([1, 0, 1, 1], F32)            # det 1 matched a crowd GT
    assign = np.array([0, 1, -1, 2], np.int32)
    gt_crowd = np.array([False, True, False])
    gt_classes = np.array([2, 1, 1], np.int32)
    cw = np.array([0.5, 2.0, 4.0], F32)               # bg, class1, class2
    r = go.loss(pred, labels, weights, assign, gt_crowd, gt_classes, cw, cfg)
    assert r['det_class'].tolist() == [2, 0, 0, 1]
    assert r['weights'].tolist() == [4.0, 0.0, 0.5, 2.0]
    ln2 = F32(np.log(2.0))
    assert abs(float(r['loss_unnormed']) - float(ln2) * 6.5) < 1e-6
    assert abs(float(r['loss_normed']) - float(ln2) * 6.5 / 4) < 1e-6
    assert r['loss'] == r['loss_unnormed']
