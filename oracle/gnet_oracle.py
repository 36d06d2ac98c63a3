"""ORACLE (test infrastructure, NOT product code): float32 numpy restatement of
the reference's Gnet forward + loss, op for op.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` leg may import this module; the product path
(`gossipnet_b200/`) never does.

Parity status: the reference itself (TensorFlow ~0.12 graph) cannot run in this
environment and ships no golden vectors. This restatement
is pinned (a) by hand-derived known-answer tests (tests/test_oracle_kat.py) and
(b) by golden vectors produced by executing the reference's OWN
`nms_net/network.py` on a numpy stand-in for the TF ops it calls
(oracle/run_reference_graph.py -> tests/golden/). See DESIGN.md §Oracle.

Every function cites the reference lines it follows (paths relative to
/root/reference). All arithmetic is float32 with every intermediate rounded
(numpy float32 arrays; python scalars are wrapped in np.float32).
"""
import numpy as np

F32 = np.float32


# --------------------------------------------------------------------------
# boxes, IoU, neighbours
# --------------------------------------------------------------------------

def xyxy_to_boxdata(a):
    """nms_net/network.py:462-472 -> (x1, y1, w, h, x2, y2, area), each [n,1]."""
    a = np.asarray(a, dtype=F32).reshape(-1, 4)
    x1, y1, x2, y2 = a[:, 0:1], a[:, 1:2], a[:, 2:3], a[:, 3:4]
    w = x2 - x1
    h = y2 - y1
    area = w * h
    return (x1, y1, w, h, x2, y2, area)


def intersection(a, b):
    """nms_net/network.py:490-511."""
    x1 = np.maximum(a[0].reshape(-1, 1), b[0].reshape(1, -1))
    y1 = np.maximum(a[1].reshape(-1, 1), b[1].reshape(1, -1))
    x2 = np.minimum(a[4].reshape(-1, 1), b[4].reshape(1, -1))
    y2 = np.minimum(a[5].reshape(-1, 1), b[5].reshape(1, -1))
    w = np.maximum(F32(0.0), x2 - x1)
    h = np.maximum(F32(0.0), y2 - y1)
    return w * h


def iou(a, b, crowd=None):
    """nms_net/network.py:474-488: inter / ((a_area + b_area) - inter); crowd
    columns use inter / a_area."""
    a_area = a[6].reshape(-1, 1)
    b_area = b[6].reshape(1, -1)
    inter = intersection(a, b)
    with np.errstate(divide='ignore', invalid='ignore'):
        union = (a_area + b_area) - inter
        out = inter / union
        if crowd is None:
            return out
        ioa = inter / a_area
    crowd = np.asarray(crowd, dtype=bool).reshape(1, -1)
    crowd = np.tile(crowd, (a_area.shape[0], 1))
    return np.where(crowd, ioa, out).astype(F32)


def class_mask_iou(det_anno_iou, det_classes, gt_classes):
    """nms_net/network.py:177-187 (multi-class only)."""
    same = det_classes.reshape(-1, 1) == gt_classes.reshape(1, -1)
    return np.where(same, det_anno_iou, F32(0.0)).astype(F32)


def neighbor_pairs(det_det_iou, thresh):
    """nms_net/network.py:192-195: tf.where(iou >= thresh) -> [P,2] int64,
    row-major order."""
    return np.argwhere(det_det_iou >= F32(thresh)).astype(np.int64)


# --------------------------------------------------------------------------
# pair features
# --------------------------------------------------------------------------

def geometry_feats(dets_boxdata, det_det_iou, det_scores, det_classes,
                   pairs, num_classes, multiplyer=1.0):
    """nms_net/network.py:411-454 (+ the multiplier at :199-200)."""
    c_idx, n_idx = pairs[:, 0], pairs[:, 1]
    n = dets_boxdata[0].shape[0]
    if num_classes > 1:
        sc = np.zeros((n, num_classes), dtype=F32)
        sc[np.arange(n), det_classes.astype(np.int64) - 1] = det_scores
    else:
        sc = det_scores.reshape(-1, 1).astype(F32)
    c_score = sc[c_idx]
    n_score = sc[n_idx]
    ious = det_det_iou[c_idx, n_idx].reshape(-1, 1)

    x1, y1, w, h = dets_boxdata[0], dets_boxdata[1], dets_boxdata[2], dets_boxdata[3]
    two = F32(2.0)
    c_w, c_h = w[c_idx], h[c_idx]
    c_scale = (c_w + c_h) / two
    c_cx = x1[c_idx] + c_w / two
    c_cy = y1[c_idx] + c_h / two
    n_w, n_h = w[n_idx], h[n_idx]
    n_cx = x1[n_idx] + n_w / two
    n_cy = y1[n_idx] + n_h / two

    x_dist = n_cx - c_cx
    y_dist = n_cy - c_cy
    l2_dist = np.sqrt(x_dist * x_dist + y_dist * y_dist) / c_scale
    x_dist = x_dist / c_scale
    y_dist = y_dist / c_scale

    log2 = F32(np.log(2.0))
    w_diff = np.log(n_w / c_w) / log2
    h_diff = np.log(n_h / c_h) / log2
    aspect_diff = (np.log(n_w / n_h) - np.log(c_w / c_h)) / log2

    out = np.concatenate([c_score, n_score, ious, x_dist, y_dist, l2_dist,
                          w_diff, h_diff, aspect_diff], axis=1).astype(F32)
    return out * F32(multiplyer)


def fc(x, params, scope, relu):
    """tf.contrib.layers.fully_connected: act(x @ W[in,out] + b)."""
    y = x @ params[scope + '/weights'] + params[scope + '/biases']
    if relu:
        y = np.maximum(y, F32(0.0))
    return y.astype(F32)


def pw_feats_fc(pw, params, cfg):
    """nms_net/network.py:324-342."""
    n_fc = cfg.gnet.num_pwfeat_fc
    for i in range(1, n_fc + 1):
        pw = fc(pw, params, 'gnet/pw_feats/fc%d' % i, relu=True)
    return pw


def segment_max(feats, seg_ids, num_segments):
    """tf.segment_max over sorted ids (network.py:387-388)."""
    starts = np.flatnonzero(np.diff(np.concatenate([[-1], seg_ids])) != 0)
    out = np.maximum.reduceat(feats, starts, axis=0)
    assert out.shape[0] == num_segments, 'a detection lost its self pair'
    return out.astype(F32)


def block(block_idx, infeats, pair_c, pair_n, pw, params, cfg):
    """nms_net/network.py:344-409."""
    g = cfg.gnet
    s = 'gnet/block%d/' % block_idx
    feats = fc(infeats, params, s + 'reduce_dim', relu=True)
    if g.neighbor_feats:
        nfeats = fc(infeats, params, s + 'reduce_dim_neighbor', relu=True)
    else:
        nfeats = feats
    c_feats = feats[pair_c]
    n_feats = nfeats[pair_n].copy()
    n_feats[pair_c == pair_n] = F32(0.0)
    x = np.concatenate([pw, c_feats, n_feats], axis=1)
    for i in range(1, g.num_block_pw_fc + 1):
        x = fc(x, params, s + 'pw_fc%d' % i, relu=True)
    x = segment_max(x, pair_c, infeats.shape[0])
    for i in range(1, g.num_block_fc):
        x = fc(x, params, s + 'fc%d' % i, relu=True)
    x = fc(x, params, s + 'fc%d' % g.num_block_fc, relu=False)
    return np.maximum(infeats + x, F32(0.0)).astype(F32)


def predict(feats, params, cfg):
    """nms_net/network.py:257-273: two LINEAR 128->128 layers, then 128->1."""
    for i in range(1, cfg.gnet.num_predict_fc):
        feats = fc(feats, params, 'gnet/predict/fc%d/fully_connected' % i, relu=False)
    return fc(feats, params, 'gnet/predict/logits/fully_connected', relu=False).reshape(-1)


# --------------------------------------------------------------------------
# loss
# --------------------------------------------------------------------------

def sigmoid_ce(x, z):
    """tf.nn.sigmoid_cross_entropy_with_logits: max(x,0) - x*z + log1p(exp(-|x|))."""
    x = x.astype(F32)
    z = z.astype(F32)
    return (np.maximum(x, F32(0.0)) - x * z
            + np.log1p(np.exp(-np.abs(x)).astype(F32))).astype(F32)


def loss(prediction, labels, weights, assignment, gt_crowd, gt_classes,
         class_weights, cfg):
    """nms_net/network.py:281-313. Returns dict with final weights and the
    three losses."""
    n = prediction.shape[0]
    if gt_crowd.shape[0] > 0:
        idx = np.maximum(assignment, 0)
        det_crowd = gt_crowd[idx]
        det_class = gt_classes.astype(np.int32)[idx]
    else:
        det_crowd = np.zeros(n, dtype=bool)
        det_class = np.zeros(n, dtype=np.int32)
    det_class = np.where((assignment >= 0) & ~det_crowd, det_class, 0)
    w = (weights * class_weights[det_class]).astype(F32)
    per = sigmoid_ce(prediction, labels) * w
    unnormed = per.sum(dtype=F32)
    normed = per.mean(dtype=F32) if n > 0 else F32(np.nan)
    base = normed if cfg.train.normalize_loss else unnormed
    return {'weights': w, 'loss_unnormed': F32(unnormed), 'loss_normed': F32(normed),
            'loss': F32(base * F32(cfg.train.loss_multiplyer)), 'det_class': det_class}


# --------------------------------------------------------------------------
# image-feature head
# --------------------------------------------------------------------------

def enlarge_windows(boxdata, padding=0.5):
    """network.py:78-87: each box grown by `padding` of its size on every side."""
    x1, y1, w, h, x2, y2, _ = boxdata
    cx = (x1 + x2) / F32(2.0)
    cy = (y1 + y2) / F32(2.0)
    nw2 = w * F32(0.5 + padding)
    nh2 = h * F32(0.5 + padding)
    return np.concatenate([cx - nw2, cy - nh2, cx + nw2, cy + nh2], axis=1).astype(F32)


def image_features(dets_boxdata, imfeats, params, cfg, stride=16):
    """network.py:103-119 (crop_windows: enlarged boxes, batch index 0, roi_pool at
    1/stride) and :223-240 (flatten -> [FC imfeat_dim relu ->] FC shortcut_dim relu).
    `imfeats` is the [1,H,W,C] map the reference gets from ResNet-101."""
    from oracle import roi_pool_oracle
    boxes = enlarge_windows(dets_boxdata)
    frcn = np.concatenate([np.zeros((boxes.shape[0], 1), dtype=F32), boxes], axis=1)
    roifeats, _ = roi_pool_oracle.roi_pool(imfeats, frcn, cfg.imfeat_crop_height,
                                           cfg.imfeat_crop_width, 1.0 / stride)
    x = roifeats.reshape(roifeats.shape[0], -1)
    scope = 'gnet/reduce_imfeats/fully_connected'
    if cfg.gnet.imfeat_dim > 0:
        x = fc(x, params, scope, relu=True)
        scope += '_1'
    return fc(x, params, scope, relu=True), roifeats, x, frcn


# --------------------------------------------------------------------------
# whole forward
# --------------------------------------------------------------------------

def gnet_forward(image, params, cfg, num_classes, matching_fn=None,
                 class_weights=None, keep_intermediates=True):
    """nms_net/network.py:148-314 for one image (dict from synthetic.make_image
    or the reference's batch spec). `matching_fn(iou, score, ignore)` is the
    DetectionMatching oracle (oracle/det_matching_oracle.py); if None or no GT
    is supplied the loss part is skipped (test.py never fetches it)."""
    g = cfg.gnet
    dets = np.asarray(image['dets'], dtype=F32)
    det_scores = np.asarray(image['det_scores'], dtype=F32)
    det_classes = np.asarray(image['det_classes'], dtype=np.int32)
    n = dets.shape[0]
    out = {}

    dets_boxdata = xyxy_to_boxdata(dets)
    det_det_iou = iou(dets_boxdata, dets_boxdata)
    pairs = neighbor_pairs(det_det_iou, g.neighbor_thresh)
    pair_c, pair_n = pairs[:, 0], pairs[:, 1]
    pw_raw = geometry_feats(dets_boxdata, det_det_iou, det_scores, det_classes,
                            pairs, num_classes, g.pw_feat_multiplyer)
    pw = pw_feats_fc(pw_raw, params, cfg) if g.num_pwfeat_fc > 0 else pw_raw

    if g.imfeats:
        feats, roifeats, det_imfeats, frcn_boxes = image_features(
            dets_boxdata, np.asarray(image['imfeats'], dtype=F32), params, cfg)
        out.update(roifeats=roifeats, det_imfeats=det_imfeats, frcn_boxes=frcn_boxes)
    else:
        feats = np.zeros((n, g.shortcut_dim), dtype=F32)
    block_feats = [feats]
    for b in range(1, g.num_blocks + 1):
        feats = block(b, feats, pair_c, pair_n, pw, params, cfg)
        block_feats.append(feats)
    prediction = predict(feats, params, cfg)

    out.update(det_det_iou=det_det_iou, neighbor_pair_idxs=pairs,
               pw_feats_raw=pw_raw, pw_feats=pw, prediction=prediction,
               num_dets=n)
    if keep_intermediates:
        out['block_feats'] = block_feats

    if 'gt_boxes' in image and image['gt_boxes'] is not None:
        gt_boxes = np.asarray(image['gt_boxes'], dtype=F32).reshape(-1, 4)
        gt_crowd = np.asarray(image['gt_crowd'], dtype=bool)
        gt_classes = np.asarray(image['gt_classes'], dtype=np.int32)
        gt_boxdata = xyxy_to_boxdata(gt_boxes)
        det_anno_iou = iou(dets_boxdata, gt_boxdata, gt_crowd)
        if num_classes > 1:
            det_anno_iou = class_mask_iou(det_anno_iou, det_classes, gt_classes)
        out['det_anno_iou'] = det_anno_iou
        if matching_fn is not None:
            labels, weights, assignment = matching_fn(det_anno_iou, prediction, gt_crowd)
            if class_weights is None:
                class_weights = np.ones(num_classes + 1, dtype=F32)
            res = loss(prediction, labels, weights, assignment, gt_crowd,
                       gt_classes, np.asarray(class_weights, dtype=F32), cfg)
            out.update(labels=labels, det_gt_matching=assignment, **res)
    return out
