"""ORACLE SUPPORT (test infrastructure): execute the reference's OWN host-side
code for the rows either side of the hot path and
store its outputs as golden fixtures.

    python oracle/run_reference_host.py          # writes tests/golden/host_logic.pkl

What runs, imported unmodified from /root/reference:
  * train.py: `_compute_ap`, `compute_aps` (AP evaluation), `LearningRate`;
  * imdb/tools.py: append_flipped, only_keep_class, drop_too_many_detections,
    get_class_counts, get_avg_batch_size, drop_no_dets;
  * imdb/coco.py: `load_detections` (FRCN detection pickle -> roidb),
    `load_image_annos` (annotation cleaning) against a minimal in-memory COCO
    stand-in object;
  * test.py: `save_dets` (roidb -> FRCN detection pickle);
  * nms_net/class_weights.py: `class_equal_weights`.
tensorflow / easydict resolve to the numpy stand-ins of oracle/tf012_numpy (none
of the functions above touches TF), pycocotools and the un-generated AnnoList_pb2 to empty stubs.  Inputs are
seeded numpy data; both inputs and outputs go into the fixture so the tests need
nothing from /root/reference.  np.round(...)+1 as a linspace count
(train.py:198) needs an int on numpy >= 1.18: the script passes through a
linspace wrapper that casts `num`, nothing else is altered.
"""
import copy
import os
import pickle
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get('GOSSIPNET_REFERENCE', '/root/reference')


def make_imdb(rs, n_images, num_classes, with_empty=True):
    classes = tuple(['__background__'] + ['c%d' % i for i in range(1, num_classes + 1)])
    roidb = []
    for i in range(n_images):
        n = int(rs.randint(0 if with_empty else 1, 40))
        g = int(rs.randint(0, 6))
        W, H = 640, 480

        def boxes(k):
            x1 = rs.uniform(0, W - 20, k)
            y1 = rs.uniform(0, H - 20, k)
            w = rs.uniform(4, 200, k)
            h = rs.uniform(4, 200, k)
            return np.stack([x1, y1, np.minimum(x1 + w, W), np.minimum(y1 + h, H)],
                            axis=1).astype(np.float32)
        roi = {'id': 1000 + i, 'width': W, 'height': H, 'filename': 'im%d.jpg' % i,
               'flipped': False,
               'gt_boxes': boxes(g), 'gt_classes': rs.randint(1, num_classes + 1, g).astype(np.int32),
               'gt_crowd': rs.uniform(0, 1, g) < 0.2}
        if n > 0 or not with_empty:
            roi.update(dets=boxes(n), det_scores=rs.uniform(0, 1, n).astype(np.float32),
                       det_classes=rs.randint(1, num_classes + 1, n).astype(np.int32))
        roidb.append(roi)
    return {'name': 'fixture', 'classes': classes,
            'class_to_ind': dict((c, i) for i, c in enumerate(classes)),
            'class_to_cat_id': dict((c, 10 * i) for i, c in enumerate(classes) if i > 0),
            'num_classes': num_classes, 'roidb': roidb}


class FakeCoco(object):
    """The three pycocotools calls the reference loader makes."""

    def __init__(self, images, anns):
        self.images = dict((im['id'], im) for im in images)
        self.anns = anns

    def loadImgs(self, i):
        return [self.images[i]]

    def getAnnIds(self, imgIds):
        return [k for k, a in enumerate(self.anns) if a['image_id'] == imgIds]

    def loadAnns(self, ids):
        return [copy.deepcopy(self.anns[k]) for k in ids]


def main(out_path):
    sys.path.insert(0, os.path.join(HERE, 'tf012_numpy'))
    sys.path.insert(0, REF)
    stub = types.ModuleType('pycocotools')
    stub.coco = types.ModuleType('pycocotools.coco')
    stub.coco.COCO = None
    sys.modules['pycocotools'] = stub
    sys.modules['pycocotools.coco'] = stub.coco
    # imdb/file_formats/pal.py imports a protobuf module the reference does not ship
    # generated (AnnoList.proto only); the PAL reader is not exercised here
    sys.modules['imdb.file_formats.AnnoList_pb2'] = types.ModuleType('imdb.file_formats.AnnoList_pb2')
    if 'scipy.misc' not in sys.modules:          # dataset.py:8 imports names scipy dropped
        import scipy.misc
        scipy.misc.imread = scipy.misc.imresize = None
    real_linspace = np.linspace
    np.linspace = lambda a, b, num=50, **kw: real_linspace(a, b, int(num), **kw)

    import tensorflow as tf
    assert 'tf012_numpy' in tf.__file__
    # network.py loads both op libraries at import time; nothing here calls them
    tf.OP_LIBRARIES['det_matching.so'] = type('MatchLib', (), {'detection_matching': None})()
    tf.OP_LIBRARIES['roi_pooling.so'] = type('RoiLib', (), {'roi_pool': None,
                                                            'roi_pool_grad': None})()
    import train as ref_train                    # /root/reference/train.py
    import test as ref_test                      # /root/reference/test.py
    import imdb.tools as ref_tools
    import imdb.coco as ref_coco
    from nms_net import cfg as ref_cfg
    from nms_net.class_weights import class_equal_weights as ref_class_weights
    assert ref_train.__file__.startswith(REF) and ref_tools.__file__.startswith(REF)

    rs = np.random.RandomState(7)
    fx = {}

    # ---- AP ----------------------------------------------------------------------
    ap_cases = []
    for n, pos, nobj in ((1, 1.0, 1), (50, 0.3, 30), (400, 0.1, 55), (400, 0.6, 200),
                         (30, 0.0, 5), (2000, 0.25, 700)):
        scores = np.sort(rs.normal(0, 2, n).astype(np.float32))[::-1].copy()
        labels = (rs.uniform(0, 1, n) < pos).astype(np.float32)
        ap_cases.append({'scores': scores, 'labels': labels, 'num_objs': nobj,
                         'ap': float(ref_train._compute_ap(scores, labels.copy(), nobj))})
    fx['compute_ap'] = ap_cases

    val_imdb = make_imdb(rs, 12, 5, with_empty=False)
    n_det = sum(r['dets'].shape[0] for r in val_imdb['roidb'])
    scores = rs.normal(0, 1, n_det).astype(np.float32)
    labels = (rs.uniform(0, 1, n_det) < 0.3).astype(np.float32)
    classes = np.concatenate([r['det_classes'] for r in val_imdb['roidb']])
    m_ap, mc_ap, cls_ap = ref_train.compute_aps(scores.copy(), classes.copy(), labels.copy(),
                                                val_imdb)
    fx['compute_aps'] = {'imdb': val_imdb, 'scores': scores, 'classes': classes, 'labels': labels,
                         'mAP': float(m_ap), 'multiclass_ap': float(mc_ap),
                         'cls_ap': [float(a) for a in cls_ap]}

    # ---- learning-rate schedule ------------------------------------------------------
    ref_cfg.train.lr_multi_step = [(5, 0.1), (9, 0.01), (12, 0.001)]
    gen = ref_train.LearningRate()
    fx['lr'] = {'steps': ref_cfg.train.lr_multi_step,
                'lrs': [gen.get_lr(it) for it in range(1, 20)]}

    # ---- imdb.tools -------------------------------------------------------------------
    base = make_imdb(rs, 10, 4)
    fx['tools_input'] = copy.deepcopy(base)
    t = {}
    t['drop_no_dets_ids'] = [r['id'] for r in ref_tools.drop_no_dets(copy.deepcopy(base)['roidb'])]
    t['append_flipped'] = ref_tools.append_flipped(copy.deepcopy(base)['roidb'])
    t['class_counts'] = ref_tools.get_class_counts(copy.deepcopy(base))
    kept = copy.deepcopy(base)
    ref_tools.only_keep_class(kept, 'c3')
    t['only_keep_class'] = kept
    cut = copy.deepcopy(base)
    ref_tools.drop_too_many_detections(cut, 7)
    t['drop_too_many'] = cut
    nodrop = copy.deepcopy(base)
    nodrop['roidb'] = ref_tools.drop_no_dets(nodrop['roidb'])
    t['avg_batch_size'] = ref_tools.get_avg_batch_size(nodrop)
    ref_cfg.train.pos_weight = 0.1
    t['class_equal_weights'] = np.asarray(ref_class_weights(copy.deepcopy(base)))
    fx['tools'] = t

    # ---- FRCN detection pickle -> roidb (coco.load_detections) --------------------------
    cat_ids = [10, 20, 30, 40]
    images = [{'id': 500 + i, 'width': 640, 'height': 480, 'file_name': 'x%d.jpg' % i}
              for i in range(6)]
    dets = [[None] * len(images) for _ in cat_ids]
    for ci in range(len(cat_ids)):
        for i in range(len(images)):
            k = int(rs.randint(0, 6))
            if i == 4:
                k = 0                                     # an image without detections
            if k == 0:
                dets[ci][i] = [] if rs.uniform() < 0.5 else np.zeros((0, 5), dtype=np.float32)
                continue
            x1 = rs.uniform(0, 600, k)
            y1 = rs.uniform(0, 440, k)
            w = rs.uniform(1, 30, k)          # some fall under det_min_size = 4
            h = rs.uniform(1, 30, k)
            dets[ci][i] = np.stack([x1, y1, x1 + w, y1 + h, rs.uniform(0, 1, k)],
                                   axis=1).astype(np.float32)
    det_file_content = (dets, [im['id'] for im in images], cat_ids)
    os.makedirs('/tmp/gn_ref_host/data', exist_ok=True)
    ref_cfg.ROOT_DIR = '/tmp/gn_ref_host'
    ref_cfg.train.detector = 'FIX'
    with open('/tmp/gn_ref_host/data/fixture_FIX.pkl', 'wb') as fp:
        pickle.dump(det_file_content, fp, protocol=2)
    cat_to_cls = dict((c, i + 1) for i, c in enumerate(cat_ids))
    coco = FakeCoco(images, [])
    fx['load_detections'] = {
        'file': det_file_content, 'cat_id_to_class_ind': cat_to_cls,
        'image_sizes': dict((im['id'], (im['width'], im['height'])) for im in images),
        'roidb': ref_coco.load_detections(coco, 'fixture', 'FIX', cat_to_cls)}

    # ---- annotations (coco.load_image_annos) ---------------------------------------------
    anns = []
    for i, im in enumerate(images):
        for _ in range(int(rs.randint(0, 5))):
            x, y = rs.uniform(-20, 620), rs.uniform(-20, 460)
            w, h = rs.uniform(-5, 120), rs.uniform(-5, 120)
            anns.append({'image_id': im['id'], 'bbox': [float(x), float(y), float(w), float(h)],
                         'area': float(max(w, 0) * max(h, 0)) * float(rs.uniform() > 0.1),
                         'iscrowd': int(rs.uniform() < 0.2),
                         'category_id': int(cat_ids[rs.randint(0, 4)])})
    coco = FakeCoco(images, anns)
    fx['annotations'] = {
        'dataset': {'images': images, 'annotations': anns,
                    'categories': [{'id': c, 'name': 'c%d' % (i + 1)} for i, c in enumerate(cat_ids)]},
        'gt_roidb': [ref_coco.load_image_annos(coco, im['id'], cat_to_cls) for im in images]}

    # ---- roidb -> FRCN detection pickle (test.save_dets) ------------------------------------
    testimdb = make_imdb(rs, 5, 3, with_empty=False)
    recs = [{'id': r['id'], 'dets': r['dets'], 'det_classes': r['det_classes'],
             'det_scores': rs.normal(0, 1, r['dets'].shape[0]).astype(np.float32)}
            for r in testimdb['roidb']]
    ref_test.save_dets(testimdb, recs, '/tmp/gn_ref_host/out.pkl')
    with open('/tmp/gn_ref_host/out.pkl', 'rb') as fp:
        fx['save_dets'] = {'imdb': testimdb, 'records': recs, 'file': pickle.load(fp)}

    with open(out_path, 'wb') as fp:
        pickle.dump(fx, fp, protocol=4)
    print('wrote', out_path, os.path.getsize(out_path), 'bytes')


if __name__ == '__main__':
    main(os.path.join(ROOT, 'tests', 'golden', 'host_logic.pkl'))
