"""CPU: the rows either side of the hot path against
golden fixtures produced by executing the reference's own train.py / test.py /
imdb code (oracle/run_reference_host.py -> tests/golden/host_logic.pkl).
AP evaluation, lr schedule, roidb tools, the FRCN detection pickle format in
both directions, COCO annotation cleaning, class weights, dataset iteration,
checkpoint state files."""
import copy
import json
import os
import pickle

import numpy as np
import pytest

from gossipnet_b200 import checkpoint, evaluation
from gossipnet_b200.imdb import coco, detections, tools
from gossipnet_b200.nms_net.config import cfg, reset_cfg
from tests.helpers import GOLDEN


@pytest.fixture(scope='module')
def fx():
    with open(os.path.join(GOLDEN, 'host_logic.pkl'), 'rb') as fp:
        return pickle.load(fp)


@pytest.fixture(autouse=True)
def _cfg():
    reset_cfg()
    yield
    reset_cfg()


def same_roi(a, b):
    assert set(a.keys()) == set(b.keys()), (sorted(a), sorted(b))
    for k in a:
        if isinstance(a[k], np.ndarray):
            assert a[k].dtype == b[k].dtype and a[k].shape == b[k].shape, k
            assert np.array_equal(a[k], b[k]), k
        else:
            assert a[k] == b[k], k


# ------------------------------------------------------------------------- AP
def test_compute_ap_matches_reference(fx):
    for case in fx['compute_ap']:
        ap = evaluation._compute_ap(case['scores'], case['labels'], case['num_objs'])
        assert ap == pytest.approx(case['ap'], abs=1e-9), case['num_objs']


def test_compute_ap_known_answers():
    # perfect ranking: precision 1 up to full recall -> AP 100
    assert evaluation._compute_ap(np.array([3., 2., 1.]), np.array([1., 1., 0.]), 2) == \
        pytest.approx(100.0)
    # one positive found last of two: precision envelope 0.5 everywhere up to recall 1
    # (recall 0 sample reads the prepended precision 1)
    ap = evaluation._compute_ap(np.array([2., 1.]), np.array([0., 1.]), 1)
    assert ap == pytest.approx((1.0 + 100 * 0.5) / 101 * 100)
    # half of the objects never detected: samples beyond recall 0.5 read 0
    ap = evaluation._compute_ap(np.array([1.]), np.array([1.]), 2)
    assert ap == pytest.approx(51.0 / 101 * 100)


def test_compute_aps_matches_reference(fx):
    c = fx['compute_aps']
    m_ap, mc_ap, cls_ap = evaluation.compute_aps(c['scores'], c['classes'], c['labels'], c['imdb'],
                                                 verbose=False)
    assert m_ap == pytest.approx(c['mAP'], abs=1e-9)
    assert mc_ap == pytest.approx(c['multiclass_ap'], abs=1e-9)
    assert np.allclose(cls_ap, c['cls_ap'], atol=1e-9)


def test_learning_rate_schedule_matches_reference(fx):
    from gossipnet_b200.trainer import LearningRate
    cfg.train.lr_multi_step = [tuple(s) for s in fx['lr']['steps']]
    gen = LearningRate()
    assert [gen.get_lr(it) for it in range(1, 20)] == fx['lr']['lrs']


# ---------------------------------------------------------------------- roidb tools
def test_roidb_tools_match_reference(fx):
    base, t = fx['tools_input'], fx['tools']
    assert [r['id'] for r in tools.drop_no_dets(copy.deepcopy(base)['roidb'])] == \
        t['drop_no_dets_ids']
    flipped = tools.append_flipped(copy.deepcopy(base)['roidb'])
    assert len(flipped) == len(t['append_flipped']) == 2 * len(base['roidb'])
    for a, b in zip(flipped, t['append_flipped']):
        same_roi(a, b)
    assert np.array_equal(tools.get_class_counts(copy.deepcopy(base)), t['class_counts'])
    kept = copy.deepcopy(base)
    tools.only_keep_class(kept, 'c3')
    for k in ('classes', 'class_to_ind', 'num_classes'):
        assert kept[k] == t['only_keep_class'][k]
    for a, b in zip(kept['roidb'], t['only_keep_class']['roidb']):
        same_roi(a, b)
    cut = copy.deepcopy(base)
    tools.drop_too_many_detections(cut, 7)
    for a, b in zip(cut['roidb'], t['drop_too_many']['roidb']):
        same_roi(a, b)
    nodrop = copy.deepcopy(base)
    nodrop['roidb'] = tools.drop_no_dets(nodrop['roidb'])
    assert tools.get_avg_batch_size(nodrop) == t['avg_batch_size']


def test_append_flipped_is_an_involution_on_boxes(fx):
    roidb = tools.drop_no_dets(copy.deepcopy(fx['tools_input'])['roidb'])
    twice = tools.append_flipped([dict(r, flipped=False) for r in
                                  tools.append_flipped(roidb)[len(roidb):]])[len(roidb):]
    for a, b in zip(roidb, twice):
        assert np.allclose(a['dets'], b['dets'], atol=1e-4)


def test_validate_boxes_rejects_degenerate_boxes():
    ok = np.array([[0, 0, 10, 10]], dtype=np.float32)
    tools.validate_boxes(ok, width=10, height=10)
    for bad in ([[0, 0, 0.5, 10]], [[-1, 0, 5, 5]], [[0, 0, 11, 5]]):
        with pytest.raises(AssertionError):
            tools.validate_boxes(np.array(bad, dtype=np.float32), width=10, height=10)


def test_class_equal_weights_matches_reference(fx):
    from gossipnet_b200.nms_net.class_weights import class_equal_weights
    cfg.train.pos_weight = 0.1
    w = class_equal_weights(copy.deepcopy(fx['tools_input']))
    assert np.allclose(w, fx['tools']['class_equal_weights'], rtol=1e-6)


# -------------------------------------------------------------- detection pickles
def test_load_detections_matches_reference(fx, tmp_path):
    c = fx['load_detections']
    path = str(tmp_path / 'dets.pkl')
    with open(path, 'wb') as fp:
        pickle.dump(c['file'], fp, protocol=2)
    roidb = detections.load_detections(path, c['cat_id_to_class_ind'], c['image_sizes'])
    assert len(roidb) == len(c['roidb'])
    for mine, ref in zip(roidb, c['roidb']):
        mine = dict((k, v) for k, v in mine.items() if k not in ('width', 'height'))
        same_roi(mine, ref)


def test_save_dets_matches_reference_and_round_trips(fx, tmp_path):
    c = fx['save_dets']
    path = str(tmp_path / 'out.pkl')
    detections.save_dets(c['imdb'], c['records'], path)
    mine = detections.read_detection_pickle(path)
    ref_dets, ref_ids, ref_cats = c['file']
    assert mine[1] == ref_ids and mine[2] == ref_cats
    for cls_mine, cls_ref in zip(mine[0], ref_dets):
        assert len(cls_mine) == len(cls_ref)
        for a, b in zip(cls_mine, cls_ref):
            assert a.shape == b.shape and a.dtype == b.dtype and np.array_equal(a, b)
    # written file -> roidb again: same detections per image, grouped by class
    cat_to_cls = dict((cat, i) for i, cat in enumerate(mine[2]) if cat != -1)
    roidb = detections.detections_to_roidb(mine[0][1:], mine[1], mine[2][1:], cat_to_cls,
                                           min_size=0)
    for rec, roi in zip(c['records'], roidb):
        order = np.argsort(rec['det_classes'], kind='stable')
        assert np.array_equal(roi['dets'], rec['dets'][order])
        assert np.array_equal(roi['det_scores'], rec['det_scores'][order])


def test_coco_annotations_match_reference(fx):
    c = fx['annotations']
    ds = c['dataset']
    an_imdb = coco.imdb_from_coco_json('fixture', ds, '/nowhere')
    assert an_imdb['num_classes'] == 4 and an_imdb['classes'][0] == '__background__'
    for roi, ref in zip(an_imdb['roidb'], c['gt_roidb']):
        for k in ('id', 'gt_boxes', 'gt_classes', 'gt_crowd'):
            if isinstance(ref[k], np.ndarray):
                assert np.array_equal(roi[k], ref[k]) and roi[k].dtype == ref[k].dtype, k
            else:
                assert roi[k] == ref[k]
        assert roi['flipped'] is False and roi['filename'].startswith('/nowhere/')


def test_load_coco_from_json_files(fx, tmp_path):
    """End to end through the files the reference reads: annotations json +
    detection pickle under cfg.ROOT_DIR/data."""
    c, d = fx['annotations'], fx['load_detections']
    root = tmp_path
    (root / 'data' / 'coco' / 'annotations').mkdir(parents=True)
    with open(str(root / 'data' / 'coco' / 'annotations' / 'instances_minival2014.json'), 'w') as fp:
        json.dump(c['dataset'], fp)
    with open(str(root / 'data' / 'coco_2014_minival_FIX.pkl'), 'wb') as fp:
        pickle.dump(d['file'], fp, protocol=2)
    cfg.ROOT_DIR = str(root)
    cfg.train.detector = 'FIX'
    an_imdb = coco.load_coco('minival', '2014')
    with_dets = [r for r in an_imdb['roidb'] if 'dets' in r]
    assert len(with_dets) == len(d['roidb'])
    for roi, ref in zip(with_dets, d['roidb']):
        assert np.array_equal(roi['dets'], ref['dets'])
        assert np.array_equal(roi['det_classes'], ref['det_classes'])
        assert 'gt_boxes' in roi and roi['width'] == 640


# --------------------------------------------------------------- imdb registry
def test_synthetic_imdb_and_preprocessing():
    import imdb as top_level
    from gossipnet_b200 import imdb as impl
    assert top_level is impl
    cfg.train.max_num_detections = 100
    train = impl.get_imdb('synthetic_train_6x250_c3', is_training=True)
    assert train['num_classes'] == 3 and len(train['roidb']) == 12      # + flipped twins
    assert all(r['dets'].shape[0] == 100 for r in train['roidb'])
    assert all(np.all(np.diff(r['det_scores']) <= 0) for r in train['roidb'])   # top-k, descending
    assert train['roidb'][6]['flipped'] and not train['roidb'][0]['flipped']
    for r in train['roidb']:
        tools.validate_boxes(r['dets'], width=r['width'], height=r['height'])
    val = impl.get_imdb('synthetic_val_4x50', is_training=False)
    assert val['num_classes'] == 1 and len(val['roidb']) == 4
    with pytest.raises(KeyError):
        impl.get_imdb('no_such_imdb', False)
    with pytest.raises(IOError):
        impl.get_imdb('coco_2014_minival', False)        # no data in this repository


def test_datasets_iterate_like_the_reference():
    from gossipnet_b200 import imdb as impl
    from gossipnet_b200.nms_net.dataset import Prefetcher, ShuffledDataset, TestDataset, load_roi
    db = impl.get_imdb('synthetic_val_5x20', is_training=False)
    ts = TestDataset(db, 1, False)
    assert len(ts) == 5
    assert [ts.next_batch()['id'] for _ in range(5)] == [r['id'] for r in db['roidb']]
    np.random.seed(3)
    sd = ShuffledDataset(db, 1, False)
    np.random.seed(3)
    perm = np.random.permutation(np.arange(5))
    epoch = [sd.next_batch() for _ in range(5)]
    assert [b['id'] for b in epoch] == [db['roidb'][i]['id'] for i in perm]
    assert epoch[0]['im_scale'] == 1.0 and epoch[0] is not db['roidb'][perm[0]]
    assert len(sd.next_batches(3)) == 3
    assert len(sd.next_batches(3)) == 3          # 2 left in the epoch -> reshuffles
    with pytest.raises(NotImplementedError):
        load_roi(True, db['roidb'][0])
    pf = Prefetcher(sd, 4, q_size=2, images_per_step=2).start()
    got = [pf.get() for _ in range(4)]
    pf.stop()
    assert all(len(g) == 2 for g in got)


# ----------------------------------------------------------------- checkpoints
def test_checkpoint_state_file_and_model_manager(tmp_path):
    d = str(tmp_path)
    assert checkpoint.get_checkpoint_state(d) is None
    checkpoint._write_state(d, 'gnet-20', ['gnet-10', 'gnet-20'])
    st = checkpoint.get_checkpoint_state(d)
    assert st == {'model_checkpoint_path': 'gnet-20',
                  'all_model_checkpoint_paths': ['gnet-10', 'gnet-20']}
    assert checkpoint.checkpoint_name('gnet', 1000) == 'gnet-1000'
    mm = checkpoint.ModelManager()
    for it, ap in ((10, 31.5), (20, 35.25), (30, 33.0)):
        path = os.path.join(d, 'gnet-%d' % it)
        open(path, 'w').close()
        mm.add(it, ap, path)
    link = os.path.join(d, 'gnet_best')
    mm.write_link_to_best(link)
    assert os.path.realpath(link) == os.path.realpath(os.path.join(d, 'gnet-20'))
    mm.add(40, 36.0, os.path.join(d, 'gnet-40'))
    mm.write_link_to_best(link)          # replaces the existing link
    assert os.readlink(link).endswith('gnet-40')
