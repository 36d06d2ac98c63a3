"""Synthetic detections / ground truth / weights (SURVEY.md §8(d) recipe).

One generator shared by tests, bench and smoke so the CUDA path and the oracle
always see the same float32 inputs. Pure numpy; no reference files are read.

Image recipe: canvas 1000x600 (reference `nms_net/config.py:21-23`),
K = max(1, N // 25) objects (centre uniform, width log-uniform in [32, 256],
aspect h/w log-uniform in [0.5, 2]); every detection jitters one object
(scale exp(N(0, .2)), centre N(0, .15) * object size), is clipped to the canvas
and keeps x2 >= x1 + 4, y2 >= y1 + 4 (`det_min_size`, `config.py:44`), so all
boxes have positive area and the IoU diagonal is exactly 1.
"""
import numpy as np

CANVAS_W = 1000.0
CANVAS_H = 600.0


def make_image(n_dets, num_classes=1, seed=42, image_index=0):
    """Return a dict of float32/int32/bool arrays named like the reference's
    batch spec (`network.py:131-146`)."""
    rs = np.random.RandomState(seed + image_index)
    n = int(n_dets)
    k = max(1, n // 25)
    ocx = rs.uniform(0.0, CANVAS_W, k)
    ocy = rs.uniform(0.0, CANVAS_H, k)
    ow = np.exp(rs.uniform(np.log(32.0), np.log(256.0), k))
    oh = ow * np.exp(rs.uniform(np.log(0.5), np.log(2.0), k))
    ocls = rs.randint(1, num_classes + 1, k)
    ocrowd = rs.uniform(0.0, 1.0, k) < 0.05

    pick = rs.randint(0, k, n)
    w = ow[pick] * np.exp(rs.normal(0.0, 0.2, n))
    h = oh[pick] * np.exp(rs.normal(0.0, 0.2, n))
    cx = ocx[pick] + rs.normal(0.0, 0.15, n) * ow[pick]
    cy = ocy[pick] + rs.normal(0.0, 0.15, n) * oh[pick]
    dets = _finish_boxes(cx, cy, w, h)

    scores = rs.uniform(0.0, 1.0, n).astype(np.float32)
    if num_classes > 1:
        classes = ocls[pick].copy()
        flip = rs.uniform(0.0, 1.0, n) < 0.10
        classes[flip] = rs.randint(1, num_classes + 1, int(flip.sum()))
    else:
        classes = np.ones(n, dtype=np.int64)

    gt = _finish_boxes(ocx, ocy, ow, oh)
    return {
        'dets': dets,
        'det_scores': scores,
        'det_classes': classes.astype(np.int32),
        'gt_boxes': gt,
        'gt_crowd': ocrowd.astype(np.bool_),
        'gt_classes': ocls.astype(np.int32),
    }


def _finish_boxes(cx, cy, w, h):
    x1 = np.clip(cx - w / 2.0, 0.0, CANVAS_W - 4.0)
    y1 = np.clip(cy - h / 2.0, 0.0, CANVAS_H - 4.0)
    x2 = np.clip(cx + w / 2.0, 0.0, CANVAS_W)
    y2 = np.clip(cy + h / 2.0, 0.0, CANVAS_H)
    x2 = np.maximum(x2, x1 + 4.0)
    y2 = np.maximum(y2, y1 + 4.0)
    return np.stack([x1, y1, x2, y2], axis=1).astype(np.float32)


def make_batch(n_images, n_dets, num_classes=1, seed=42, first_index=0):
    return [make_image(n_dets, num_classes, seed, first_index + i)
            for i in range(n_images)]
