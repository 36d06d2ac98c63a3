"""GPU: the drivers around the hot path end to end on
synthetic imdbs: train.py loop with validation, checkpoints, best-model link and
resume; test.py rescoring into the Fast R-CNN pickle; batched validation equals
the reference's per-image loop."""
import os
import sys

import numpy as np
import pytest
import torch

from gossipnet_b200 import evaluation, imdb
from gossipnet_b200.imdb import detections
from gossipnet_b200.nms_net.config import cfg
from gossipnet_b200.nms_net.network import Gnet
from tests.helpers import load_experiment

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def small_cfg(blocks=2):
    load_experiment('coco_person', num_blocks=blocks)
    cfg.train.imdb = 'synthetic_train_4x60'
    cfg.train.val_imdb = 'synthetic_val_3x50'
    cfg.test.imdb = 'synthetic_val_3x50'
    cfg.train.lr_multi_step = [(4, 0.001), (100, 0.0001)]
    cfg.train.display_iter = 2


def import_driver(name):
    sys.path.insert(0, ROOT)
    try:
        return __import__(name)
    finally:
        sys.path.remove(ROOT)


def test_train_driver_checkpoints_validation_and_resume(tmp_path, monkeypatch, capsys):
    monkeypatch.chdir(tmp_path)
    small_cfg()
    cfg.train.num_iter, cfg.train.save_iter, cfg.train.val_iter = 6, 2, 4
    train = import_driver('train')
    trainer = train.train(resume=False, visualize=False, images_per_step=2)
    out = capsys.readouterr().out
    assert 'validation pass:   mAP' in out and 'training finished' in out
    assert 'opt loss' in out
    for it in (2, 4, 6):
        assert os.path.exists('gnet-%d' % it)
    assert os.readlink('gnet_best').endswith('gnet-4')     # the only validated model
    state = open('checkpoint').read()
    assert 'model_checkpoint_path: "gnet-6"' in state and state.count('all_model') == 3
    params_after_6 = trainer.eng.flat.clone()
    steps_after_6 = trainer.global_step

    # resume: continues at iteration 7 with the saved parameters and Adam slots
    cfg.train.num_iter = 8
    resumed = train.train(resume=True, visualize=False, images_per_step=2)
    out = capsys.readouterr().out
    assert 'resuming at iteration 7' in out
    assert resumed.global_step == steps_after_6 + 2
    assert os.path.exists('gnet-8')
    assert not torch.equal(resumed.eng.flat, params_after_6)

    # an uninterrupted 8-iteration run lands on the same parameters (same data order:
    # the permutation stream is re-seeded, so replay it)
    for f in os.listdir('.'):
        os.remove(f)
    straight = train.train(resume=False, visualize=False, images_per_step=2)
    # resume re-seeds the permutation stream, so iterations 7-8 saw other images than in
    # the straight run: parameters differ slightly but stay close after two Adam steps
    rel = float((straight.eng.flat - resumed.eng.flat).abs().max())
    assert rel < 5e-3


def test_test_driver_writes_frcn_pickle(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    small_cfg()
    test_imdb = imdb.get_imdb(cfg.test.imdb, is_training=False)
    test = import_driver('test')
    with pytest.raises(ValueError):        # no model configured: refuse, like the reference
        test.test_run(test_imdb, images_per_call=2)
    dets = test.test_run(test_imdb, images_per_call=2, allow_random_init=True)   # 3 images: chunks of 2 + 1
    assert [d['id'] for d in dets] == [r['id'] for r in test_imdb['roidb']]
    net = Gnet(1, reuse=True)
    for d, roi in zip(dets, test_imdb['roidb']):
        single = net({k: roi[k] for k in ('dets', 'det_scores', 'det_classes')}).cpu().numpy()
        assert np.array_equal(d['det_scores'], single)
    detections.save_dets(test_imdb, dets, 'out.pkl')
    frcn, image_ids, cat_ids = detections.read_detection_pickle('out.pkl')
    assert cat_ids == [-1, 1] and len(frcn) == 2 and len(frcn[1]) == 3
    assert all(x.shape == (0, 5) for x in frcn[0])
    assert frcn[1][0].shape == (50, 5)
    assert np.array_equal(frcn[1][2][:, 4], dets[2]['det_scores'])


def test_batched_validation_equals_per_image_loop():
    small_cfg()
    cfg.train.val_imdb = 'synthetic_val_5x80_c3'
    cfg.train.only_class = ''                # keep all three classes
    val_imdb = imdb.get_imdb(cfg.train.val_imdb, is_training=False)
    net = Gnet(3)
    batched = evaluation.collect_val_outputs(net, val_imdb, images_per_call=4)
    scores, classes, labels = [], [], []
    for roi in val_imdb['roidb']:            # train.py:140-154, one image per run
        net(roi)
        keep = (net.weights > 0).cpu().numpy()
        scores.append(net.prediction.cpu().numpy()[keep])
        labels.append(net.labels.cpu().numpy()[keep])
        classes.append(roi['det_classes'][keep])
    assert np.array_equal(batched[0], np.concatenate(scores))
    assert np.array_equal(batched[1], np.concatenate(classes))
    assert np.array_equal(batched[2], np.concatenate(labels))
    m_ap, mc_ap, cls_ap = evaluation.val_run(net, val_imdb, images_per_call=4, verbose=False)
    assert 0.0 <= mc_ap <= 100.0 and len(cls_ap) == 3


def test_session_graphs_repeated_shapes_and_runs_new_shapes_eagerly():
    """InferenceSession: a batch shape that comes back is replayed as a CUDA graph (captured
    the second time it is seen), one-off shapes run eagerly through grow-only staging - and
    every path returns exactly what Gnet.run_batch computes."""
    from gossipnet_b200 import synthetic
    from gossipnet_b200.session import InferenceSession
    load_experiment('coco_person', num_blocks=2)
    net = Gnet(1)
    sess = InferenceSession(net)

    def batch(sizes, first):
        imgs = [synthetic.make_image(n, 1, image_index=first + i) for i, n in enumerate(sizes)]
        off = np.zeros(len(sizes) + 1, np.int32)
        np.cumsum(sizes, out=off[1:])
        cat = lambda k, dt: np.concatenate([im[k] for im in imgs]).astype(dt)
        want = net.run_batch([{k: im[k] for k in ('dets', 'det_scores', 'det_classes')}
                              for im in imgs])['prediction'].cpu().numpy().copy()
        return (cat('dets', np.float32), cat('det_scores', np.float32),
                cat('det_classes', np.int32), off), want

    shapes = [[120, 80], [300], [120, 80], [50, 60, 70], [120, 80], [300], [640, 1], [120, 80]]
    graphed = []
    for i, sizes in enumerate(shapes):
        args, want = batch(sizes, first=10 * i)          # same shape, different boxes each time
        got = sess.run(*args)
        assert np.array_equal(got, want), (i, sizes)
        graphed.append(bool(sess._graph))
    # [120,80] is captured at its 2nd appearance, [300] at its 2nd; the one-off shapes never
    assert graphed == [False, False, True, False, True, True, False, True]
    assert len(sess._states) == 2
    assert sess.run(np.zeros((0, 4), np.float32), np.zeros(0, np.float32), np.zeros(0, np.int32),
                    np.zeros(1, np.int32)).shape == (0,)


def test_pipelined_session_equals_blocking_calls():
    """InferenceSession.run_pipelined (one batch ahead, two staging slots) returns, in order,
    exactly what the blocking run() returns - for repeated shapes (captured graphs), a shape
    change in the middle, an empty batch, and a batch denser than the workspace."""
    from gossipnet_b200 import synthetic
    from gossipnet_b200.session import InferenceSession
    small_cfg()
    net = Gnet(1)

    def batch(sizes, first):
        imgs = [synthetic.make_image(n, 1, image_index=first + i) for i, n in enumerate(sizes)]
        off = np.zeros(len(sizes) + 1, np.int32)
        np.cumsum(sizes, out=off[1:])
        return (np.concatenate([i['dets'] for i in imgs]), np.concatenate([i['det_scores'] for i in imgs]),
                np.concatenate([i['det_classes'] for i in imgs]), off)

    seq = [batch([120, 80], 0), batch([120, 80], 2), batch([120, 80], 4), batch([50], 6),
           batch([120, 80], 7), (np.zeros((0, 4), np.float32), np.zeros(0, np.float32),
                                 np.zeros(0, np.int32), np.zeros(1, np.int32)),
           batch([120, 80], 9), batch([900, 900], 11), batch([120, 80], 13)]
    ref_sess = InferenceSession(net, use_graph=False)
    want = [ref_sess.run(*b).copy() for b in seq]
    # a second engine with the same weights: its workspace has to grow on the dense batch
    sess = InferenceSession(Gnet(1, params=net.engine.flat.cpu().numpy()), graph_after=1)
    got = [o.copy() for o in sess.run_pipelined(seq)]
    assert len(got) == len(want)
    for a, b in zip(got, want):
        assert np.array_equal(a, b)
    got2 = [o.copy() for o in sess.run_pipelined(seq)]       # second pass: everything captured
    for a, b in zip(got2, want):
        assert np.array_equal(a, b)
