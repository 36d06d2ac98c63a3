"""COCO -> imdb (reference imdb/coco.py:17-205), without pycocotools: the
annotation files are plain JSON, and the loader only needs the category list,
the image list and each image's annotations in file order (what
`COCO.getCatIds / getImgIds / getAnnIds(imgIds=...)` return).

imdb = {name, classes ('__background__', ...), class_to_ind, class_to_cat_id,
        num_classes, roidb}; roidb entry = {id, width, height, filename, flipped,
        [gt_boxes, gt_classes, gt_crowd], [dets, det_scores, det_classes]}.
Detections come from `data/<imdb name>_<cfg.train.detector>.pkl`
(gossipnet_b200.imdb.detections).  There is no COCO data in this repository;
tests build a miniature annotation file and detection pickle.
"""
import json
import os.path
import pickle

import numpy as np

from gossipnet_b200.imdb import detections
from gossipnet_b200.imdb.tools import validate_boxes
from gossipnet_b200.nms_net.config import cfg

IMAGE_DIRS = {
    'coco_2014_train': 'train2014', 'coco_2014_debug': 'train2014',
    'coco_2014_val': 'val2014', 'coco_2014_minival': 'val2014',
    'coco_2014_valminusminival': 'val2014',
}


def clean_annotation_boxes(objs, width, height):
    """coco.py:191-201: clip xywh boxes to the canvas, keep positive-area ones."""
    kept = []
    for obj in objs:
        x1 = max(0, obj['bbox'][0])
        y1 = max(0, obj['bbox'][1])
        x2 = min(width, x1 + max(0, obj['bbox'][2]))
        y2 = min(height, y1 + max(0, obj['bbox'][3]))
        if obj['area'] > 0 and x2 >= x1 and y2 >= y1:
            kept.append((obj, [x1, y1, x2, y2]))
    return kept


def image_annotations(im_info, objs, cat_id_to_class_ind, min_size=None):
    """coco.py:152-188 for one image."""
    min_size = cfg.train.det_min_size if min_size is None else min_size
    kept = clean_annotation_boxes(objs, im_info['width'], im_info['height'])
    boxes = np.zeros((len(kept), 4), dtype=np.float32)
    classes = np.zeros((len(kept),), dtype=np.int32)
    crowd = np.zeros((len(kept),), dtype=np.bool_)
    for i, (obj, box) in enumerate(kept):
        boxes[i, :] = box
        crowd[i] = obj['iscrowd']
        classes[i] = cat_id_to_class_ind[obj['category_id']]
    big = np.logical_and(boxes[:, 2] - boxes[:, 0] >= min_size,
                         boxes[:, 3] - boxes[:, 1] >= min_size)
    boxes, classes, crowd = boxes[big, :], classes[big], crowd[big]
    validate_boxes(boxes, width=im_info['width'], height=im_info['height'])
    return {'id': im_info['id'], 'gt_boxes': boxes, 'gt_classes': classes, 'gt_crowd': crowd}


def merge_roidbs(roidb_a, roidb_b):
    """coco.py:122-129: update entries of a with the entry of b of the same id."""
    assert len(roidb_a) >= len(roidb_b)
    by_id = dict((r['id'], r) for r in roidb_a)
    for rb in roidb_b:
        by_id[rb['id']].update(rb)
    return roidb_a


def imdb_from_coco_json(name, dataset, image_dir, detection_file=None):
    cats = dataset['categories']
    classes = tuple(['__background__'] + [c['name'] for c in cats])
    class_to_ind = dict((c, i) for i, c in enumerate(classes))
    class_to_cat_id = dict((c['name'], c['id']) for c in cats)
    cat_id_to_class_ind = dict((class_to_cat_id[c], class_to_ind[c]) for c in classes[1:])
    roidb = [{'id': im['id'], 'width': im['width'], 'height': im['height'],
              'filename': os.path.join(image_dir, im['file_name']), 'flipped': False}
             for im in dataset['images']]
    if 'annotations' in dataset:
        per_image = {}
        for ann in dataset['annotations']:
            per_image.setdefault(ann['image_id'], []).append(ann)
        gt = [image_annotations(im, per_image.get(im['id'], []), cat_id_to_class_ind)
              for im in dataset['images']]
        roidb = merge_roidbs(roidb, gt)
    if detection_file is not None:
        sizes = dict((im['id'], (im['width'], im['height'])) for im in dataset['images'])
        det_roidb = detections.load_detections(detection_file, cat_id_to_class_ind, sizes)
        for r in det_roidb:      # width / height already live in the image entry
            r.pop('width'), r.pop('height')
        roidb = merge_roidbs(roidb, det_roidb)
    return {'name': name, 'classes': classes, 'class_to_ind': class_to_ind,
            'class_to_cat_id': class_to_cat_id, 'num_classes': len(classes) - 1, 'roidb': roidb}


def load_coco(split, year):
    """coco.py:17-70 (same files, same cache)."""
    name = 'coco_{}_{}'.format(year, split)
    cache_file = os.path.join(cfg.ROOT_DIR, 'data', 'cache', '{}.pkl'.format(name))
    if os.path.exists(cache_file):
        with open(cache_file, 'rb') as fp:
            return pickle.load(fp)
    ann_dir = os.path.join(cfg.ROOT_DIR, 'data', 'coco', 'annotations')
    ann_file = os.path.join(ann_dir, 'instances_{}{}.json'.format(split, year))
    if not os.path.exists(ann_file):
        ann_file = os.path.join(ann_dir, 'image_info_{}{}.json'.format(split, year))
    if not os.path.exists(ann_file):
        raise IOError('no COCO annotations for {} under {} (this repository ships no data; '
                      'use a synthetic_* imdb or place the files there)'.format(name, ann_dir))
    with open(ann_file) as fp:
        dataset = json.load(fp)
    image_dir = os.path.join(cfg.ROOT_DIR, 'data', 'coco', 'images',
                             IMAGE_DIRS.get(name, split + year))
    det_file = os.path.join(cfg.ROOT_DIR, 'data', '{}_{}.pkl'.format(name, cfg.train.detector))
    return imdb_from_coco_json(name, dataset, image_dir, det_file)
