"""Validation pass of the reference (train.py:133-206) on the B200 surface.

`val_run` feeds the validation roidb through the Gnet in batches of images
(one engine call per batch instead of one sess.run per image), drops the
detections matched to crowd annotations (`weights > 0`), and scores the rest
with `compute_aps`: detections sorted by descending new score, running
precision made monotone from the right, precision sampled at the 101 recall
points 0, 0.01, ..., 1 (`np.searchsorted(..., side='left')`), AP = mean * 100.
The arithmetic is the reference's (float32 cumulative counts); it is host-side
numpy exactly like there.
"""
import numpy as np


def _compute_ap(scores, labels, num_objs):
    """train.py:186-204.  `scores` must already be sorted descending."""
    labels = np.asarray(labels)
    fp = np.cumsum((labels == 0).astype(np.int32)).astype(np.float32)
    tp = np.cumsum((labels == 1).astype(np.int32)).astype(np.float32)
    with np.errstate(divide='ignore', invalid='ignore'):
        recall = tp / num_objs
        precision = tp / (fp + tp)
    # envelope: precision[i] = max(precision[i:])
    if precision.size:
        precision = np.maximum.accumulate(precision[::-1])[::-1]
    last = recall[-1] if recall.size else 0.0
    recall = np.concatenate(([0], recall, [last, 2]), axis=0)
    precision = np.concatenate(([1], precision, [0, 0]), axis=0)
    points = np.linspace(0.0, 1.0, 101, endpoint=True)
    inds = np.searchsorted(recall, points, side='left')
    return np.average(precision[inds]) * 100


def compute_aps(scores, classes, labels, val_imdb, verbose=True):
    """train.py:161-184 -> (mAP over the classes present among the detections,
    class-agnostic AP, [per-class AP])."""
    order = np.argsort(-scores)
    scores, labels, classes = scores[order], labels[order], classes[order]
    roidb = val_imdb['roidb']
    num_objs = sum(np.sum(np.logical_not(roi['gt_crowd'])) for roi in roidb)
    multiclass_ap = _compute_ap(scores, labels, num_objs)
    present = np.unique(classes)
    if verbose:
        print(present)
    cls_ap = []
    for c in present:
        m = classes == c
        c_objs = sum(np.sum(np.logical_and(np.logical_not(roi['gt_crowd']),
                                           roi['gt_classes'] == c)) for roi in roidb)
        cls_ap.append(_compute_ap(scores[m], labels[m], c_objs))
    return np.mean(cls_ap), multiclass_ap, cls_ap


def collect_val_outputs(net, val_imdb, images_per_call=64):
    """-> (scores, classes, labels) of every non-ignored detection of the imdb
    (train.py:133-158), one batched engine call per `images_per_call` images."""
    from gossipnet_b200.nms_net.dataset import load_roi
    rois = [load_roi(False, roi) for roi in val_imdb['roidb']
            if 'dets' in roi and roi['dets'].size > 0]
    all_scores, all_labels, all_classes = [], [], []
    for i in range(0, len(rois), images_per_call):
        chunk = rois[i:i + images_per_call]
        res = net.run_batch(chunk)
        keep = (res['weights'] > 0.0).cpu().numpy()
        all_scores.append(res['prediction'].cpu().numpy()[keep])
        all_labels.append(res['labels'].cpu().numpy()[keep])
        all_classes.append(np.concatenate([r['det_classes'] for r in chunk])[keep])
    return (np.concatenate(all_scores), np.concatenate(all_classes), np.concatenate(all_labels))


def val_run(net, val_imdb, images_per_call=64, verbose=True):
    """train.py:133-158.  The reference's `sess` argument has no equivalent."""
    scores, classes, labels = collect_val_outputs(net, val_imdb, images_per_call)
    return compute_aps(scores, classes, labels, val_imdb, verbose=verbose)
