"""roidb utilities on the consumer side of the hot path (reference
imdb/tools.py:8-126).  A roidb is a list of per-image dicts with any of
`dets[n,4] det_scores[n] det_classes[n] gt_boxes[g,4] gt_classes[g] gt_crowd[g]
width height flipped id filename`.  Same names, argument meaning and in-place
behaviour as the reference; written against numpy only."""
import numpy as np


def _mirror_x(boxes, width):
    out = np.array(boxes, copy=True)
    out[:, 0] = width - boxes[:, 2]
    out[:, 2] = width - boxes[:, 0]
    return out


def append_flipped(roidb):
    """roidb + its horizontal mirror images (tools.py:8-28): x1' = W - x2,
    x2' = W - x1, `flipped` = True, everything else shared."""
    mirrored = []
    for roi in roidb:
        twin = dict(roi)
        twin['flipped'] = True
        for key in ('dets', 'gt_boxes'):
            if key in roi:
                twin[key] = _mirror_x(roi[key], roi['width'])
        mirrored.append(twin)
    return roidb + mirrored


def drop_no_dets(roidb):
    """tools.py:31-33."""
    return [roi for roi in roidb if 'dets' in roi and roi['dets'].size > 0]


def drop_no_gt(roidb):
    """tools.py:36-38."""
    return [roi for roi in roidb if 'gt_boxes' in roi]


def only_keep_class(imdb, class_name):
    """Single-class view of a multi-class imdb, in place (tools.py:41-63): the
    kept class becomes class 1, `num_classes` 1."""
    keep = imdb['class_to_ind'][class_name]
    imdb['classes'] = (imdb['classes'][0], imdb['classes'][keep])
    imdb['class_to_ind'] = dict((name, i) for i, name in enumerate(imdb['classes']))
    imdb['num_classes'] = 1
    for roi in imdb['roidb']:
        if 'gt_classes' in roi:
            m = roi['gt_classes'] == keep
            roi['gt_boxes'] = roi['gt_boxes'][m, :].copy()
            roi['gt_crowd'] = roi['gt_crowd'][m].copy()
            roi['gt_classes'] = np.ones(int(m.sum()), dtype=roi['gt_classes'].dtype)
        if 'det_classes' in roi:
            m = roi['det_classes'] == keep
            roi['dets'] = roi['dets'][m, :].copy()
            roi['det_scores'] = roi['det_scores'][m].copy()
            roi['det_classes'] = np.ones(int(m.sum()), dtype=roi['det_classes'].dtype)
            validate_boxes(roi['dets'], width=roi['width'], height=roi['height'])


def drop_too_many_detections(imdb, max_num_detections):
    """Keep the `max_num_detections` highest-scoring detections of every image,
    in descending score order (tools.py:66-74: argsort ascending, reversed) -
    the step immediately before the hot path."""
    for roi in imdb['roidb']:
        if 'det_classes' not in roi:
            continue
        top = np.argsort(roi['det_scores'])[::-1][:max_num_detections]
        roi['dets'] = roi['dets'][top, :]
        roi['det_scores'] = roi['det_scores'][top]
        roi['det_classes'] = roi['det_classes'][top]


def stats(imdb):
    """(images, detections, crowd annotations, non-crowd annotations)."""
    n_det = n_crowd = n_anno = 0
    for roi in imdb['roidb']:
        if 'gt_crowd' in roi:
            c = int(np.sum(roi['gt_crowd']))
            n_crowd += c
            n_anno += roi['gt_boxes'].shape[0] - c
        if 'dets' in roi:
            n_det += roi['dets'].shape[0]
    return len(imdb['roidb']), n_det, n_crowd, n_anno


def print_stats(imdb):
    """tools.py:77-90 (same line)."""
    print('{:d} images: {:d} detections, {:d} crowd annotations, '
          '{:d} non-crowd annotations'.format(*stats(imdb)))


def get_avg_batch_size(imdb):
    """tools.py:93-96."""
    total = sum(roi['dets'].shape[0] for roi in imdb['roidb'] if 'dets' in roi)
    return total / len(imdb['roidb'])


def validate_boxes(boxes, width=0, height=0):
    """tools.py:99-111: boxes lie on the canvas and are at least 1 px wide/high
    (this is what guarantees area > 0, i.e. an IoU diagonal of exactly 1)."""
    x1, y1, x2, y2 = boxes[:, 0], boxes[:, 1], boxes[:, 2], boxes[:, 3]
    assert (x1 >= 0).all() and (y1 >= 0).all()
    assert (x2 >= x1 + 1).all() and (y2 >= y1 + 1).all()
    assert (x2 <= width).all() and (y2 <= height).all()


def get_class_counts(imdb):
    """Samples per class for the loss weights (tools.py:114-126): every count
    starts at 1; each GT adds to its class; detections beyond the number of GTs
    of the image count as background."""
    freq = np.ones((imdb['num_classes'] + 1,), dtype=np.int64)
    for roi in imdb['roidb']:
        n_pos = 0
        if 'gt_classes' in roi:
            n_pos = roi['gt_classes'].size
            np.add.at(freq, np.asarray(roi['gt_classes'], dtype=np.int64), 1)
        if 'det_classes' in roi:
            freq[0] += max(0, roi['det_classes'].size - n_pos)
    return freq
