#!/usr/bin/env python
"""Inference driver: the reference's test.py (`:42-136`) on the B200 path.

    python test.py OUTFILE -c conf.yaml -m gnet-10000 [-s synthetic_val_64x1000]

Same flags; restores the checkpoint, rescores every image of the test imdb and
writes the detections in the Fast R-CNN pickle layout.  The reference runs one
sess.run per image and times feed + run + fetch (test.py:69-71); here
`--images-per-call` images (default 64) go through ONE host-buffer session call
(pinned H2D of the boxes/scores/classes, the whole forward as a CUDA graph per
batch shape, D2H of the logits) and the same timer wraps that call.
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import imdb  # noqa: E402
from imdb.detections import save_dets  # noqa: E402
from nms_net import cfg, tools  # noqa: E402
from nms_net.config import cfg_from_file  # noqa: E402
from nms_net.dataset import load_roi  # noqa: E402
from nms_net.network import Gnet  # noqa: E402
from gossipnet_b200.checkpoint import load_variables  # noqa: E402
from gossipnet_b200.session import InferenceSession  # noqa: E402


def test_run(test_imdb, images_per_call=64, model=None, allow_random_init=False):
    """test.py:42-83 -> [{'id', 'dets', 'det_classes', 'det_scores'}, ...].  Like the
    reference (restorer.restore(sess, cfg.test_model), test.py:52-54) this needs a trained
    model: without one it raises instead of rescoring with random weights
    (`allow_random_init` is for tests of the plumbing)."""
    roidb = test_imdb['roidb']
    batch_spec = Gnet.get_batch_spec(num_classes=test_imdb['num_classes'], is_training=False)
    need_image = 'image' in batch_spec

    net = Gnet(num_classes=test_imdb['num_classes'])
    model = cfg.get('test_model') if model is None else model
    if model:
        net.load_state_dict(load_variables(model))
    elif not allow_random_init:
        raise ValueError('no model to test: pass -m / --model or set cfg.test_model')
    sess = InferenceSession(net)

    rois = [load_roi(need_image, roi) for roi in roidb
            if 'dets' in roi and roi['dets'].size > 0]
    output_detections = []
    forward_timer = tools.Timer()
    num_dets = num_images = 0
    def chunks():
        for i in range(0, len(rois), images_per_call):
            chunk = rois[i:i + images_per_call]
            sizes = [r['dets'].shape[0] for r in chunk]
            off = np.zeros(len(chunk) + 1, dtype=np.int32)
            np.cumsum(sizes, out=off[1:])
            yield chunk, off, (np.concatenate([r['dets'] for r in chunk]).astype(np.float32),
                               np.concatenate([r['det_scores'] for r in chunk]).astype(np.float32),
                               np.concatenate([r['det_classes'] for r in chunk]).astype(np.int32), off)

    # the session runs one chunk ahead: while the GPU rescored chunk i the host has already
    # staged chunk i+1 (InferenceSession.run_pipelined); results come back in order
    meta = []

    def inputs():
        for chunk, off, args in chunks():
            meta.append((chunk, off))
            yield args

    forward_timer.tic()
    for k_chunk, new_scores in enumerate(sess.run_pipelined(inputs())):
        chunk, off = meta[k_chunk]
        new_scores = new_scores.copy()
        for k, roi in enumerate(chunk):
            output_detections.append({
                'id': roi['id'],
                'dets': roi['dets'] / roi['im_scale'],
                'det_classes': roi['det_classes'],
                'det_scores': new_scores[off[k]:off[k + 1]],
            })
        num_dets += int(off[-1])
        num_images += len(chunk)
    forward_timer.toc()
    if num_images:
        print('{:.6f}s per image with {:.1f} detections per image'.format(
            forward_timer.total_time / num_images, num_dets / num_images))
    return output_detections


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument('outfile', help='detection file output')
    parser.add_argument('-c', '--config', default='conf.yaml')
    parser.add_argument('-m', '--model', default=None)
    parser.add_argument('-s', '--imdb', default=None)
    parser.add_argument('--images-per-call', type=int, default=64)
    args, unparsed = parser.parse_known_args()

    cfg_from_file(args.config)
    if args.model is not None:
        cfg.test_model = args.model
    if args.imdb is not None:
        cfg.test.imdb = args.imdb

    test_imdb = imdb.get_imdb(cfg.test.imdb, is_training=False)
    dets = test_run(test_imdb, args.images_per_call)
    save_dets(test_imdb, dets, args.outfile)


if __name__ == '__main__':
    main()
