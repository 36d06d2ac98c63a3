"""The on-disk detection format either side of the hot path: Fast R-CNN style
pickles `(dets, image_ids, cat_ids)` with `dets[class_index][image_index]` a
float array [n, 5] = (x1, y1, x2, y2, score).

  * input : reference imdb/coco.py:73-119 (`load_detections`) - per image the
            classes are concatenated in `cat_ids` order, detections smaller than
            `cfg.train.det_min_size` are dropped, boxes are validated;
  * output: reference test.py:86-111 (`save_dets`) - rescored detections back to
            the same layout, pickle protocol 2.
"""
import pickle

import numpy as np

from gossipnet_b200.imdb.tools import validate_boxes
from gossipnet_b200.nms_net.config import cfg


def read_detection_pickle(filename):
    with open(filename, 'rb') as fp:
        dets, image_ids, cat_ids = pickle.load(fp, encoding='latin1')
    return dets, image_ids, cat_ids


def detections_to_roidb(dets, image_ids, cat_ids, cat_id_to_class_ind, image_sizes=None,
                        min_size=None):
    """-> list of dicts {id, dets[n,4], det_scores[n], det_classes[n] int32}.
    Images without a single detection are skipped (coco.py:95-96);
    `image_sizes[id] = (width, height)` enables the box validation and adds
    width / height to the entry."""
    min_size = cfg.train.det_min_size if min_size is None else min_size
    roidb = []
    for i, imid in enumerate(image_ids):
        boxes, scores, classes = [], [], []
        for ci, cat_id in enumerate(cat_ids):
            d = dets[ci][i]
            if isinstance(d, list):
                if len(d) == 0:
                    continue
                d = np.asarray(d)
            if d.size == 0:
                continue
            boxes.append(d[:, :4])
            scores.append(d[:, 4])
            classes.append(np.full((d.shape[0],), cat_id_to_class_ind[cat_id], dtype=np.int32))
        if not classes:
            continue
        boxes = np.concatenate(boxes, axis=0)
        scores = np.concatenate(scores, axis=0)
        classes = np.concatenate(classes, axis=0)
        big = np.logical_and(boxes[:, 2] - boxes[:, 0] >= min_size,
                             boxes[:, 3] - boxes[:, 1] >= min_size)
        entry = {'id': imid, 'dets': boxes[big, :], 'det_scores': scores[big],
                 'det_classes': classes[big]}
        if image_sizes is not None:
            w, h = image_sizes[imid]
            validate_boxes(entry['dets'], width=w, height=h)
            entry['width'], entry['height'] = w, h
        roidb.append(entry)
    return roidb


def load_detections(filename, cat_id_to_class_ind, image_sizes=None, min_size=None):
    return detections_to_roidb(*read_detection_pickle(filename),
                               cat_id_to_class_ind=cat_id_to_class_ind,
                               image_sizes=image_sizes, min_size=min_size)


def dets_to_frcn(testimdb, dets_as_dicts):
    """[{id, dets[n,4], det_classes[n], det_scores[n]}, ...] ->
    (dets[class][image] -> [k,5], image_ids, cat_ids)   (test.py:86-108).
    Class index 0 is the background entry: cat id -1, always empty unless some
    detection carries class 0."""
    cat_ids = [testimdb['class_to_cat_id'].get(name, -1) for name in testimdb['classes']]
    out = [[] for _ in cat_ids]
    image_ids = []
    for rec in dets_as_dicts:
        image_ids.append(rec['id'])
        cls = np.asarray(rec['det_classes'])
        for ci in range(len(cat_ids)):
            m = cls == ci
            if m.any():
                rows = np.concatenate((rec['dets'][m, :], rec['det_scores'][m][:, None]), axis=1)
            else:
                rows = np.zeros((0, 5), dtype=np.float32)
            out[ci].append(rows)
    return out, image_ids, cat_ids


def save_dets(testimdb, dets_as_dicts, output_file):
    """test.py:86-111: pickle protocol 2 (readable from Python 2.7)."""
    with open(output_file, 'wb') as fp:
        pickle.dump(dets_to_frcn(testimdb, dets_as_dicts), fp, protocol=2)
