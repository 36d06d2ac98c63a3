#!/usr/bin/env python
"""Training driver: the reference's train.py (CLI `:402-416`, loop `:222-354`)
on the B200 path.

    python train.py -c experiments/coco_person/conf.yaml [-r] [--images-per-step K]
    python -m torch.distributed.run --nproc-per-node 8 train.py -c ... --images-per-step 64

Same flags and conf.yaml keys; same loop: lr_multi_step schedule, display every
`display_iter`, validation + checkpoint + `gnet_best` symlink every `val_iter`,
checkpoint every `save_iter` and at the end, `-r` resumes from the newest
checkpoint in the working directory at its iteration + 1.  Differences: the
tf.Session / FIFOQueue machinery is a `Trainer` step plus a prefetch thread;
`--images-per-step` (default 1 = the reference) trains on several images per
step, sharded over the ranks of a torchrun launch with ONE NCCL all-reduce of
the flat gradient; `-v` (matplotlib over loaded images) is not
available.  With no COCO data in the repository, point `train.imdb` /
`train.val_imdb` at `synthetic_*` imdbs.
"""
import argparse
import os
import sys
from datetime import datetime
from pprint import pprint

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import imdb  # noqa: E402
from nms_net import cfg  # noqa: E402
from nms_net.class_weights import class_equal_weights  # noqa: E402
from nms_net.config import cfg_from_file  # noqa: E402
from nms_net.dataset import Prefetcher, ShuffledDataset  # noqa: E402
from nms_net.network import Gnet  # noqa: E402
from gossipnet_b200 import evaluation, parallel  # noqa: E402
from gossipnet_b200.checkpoint import ModelManager, Saver  # noqa: E402
from gossipnet_b200.trainer import LearningRate, Trainer  # noqa: E402


class SmoothedLosses(object):
    """tf.train.ExponentialMovingAverage(decay=0.7) over loss tensors
    (train.py:263-272): zero-initialised shadows, no debiasing."""

    def __init__(self, decay=0.7):
        self.decay, self.avg = decay, {}

    def update(self, **values):
        for k, v in values.items():
            self.avg[k] = self.decay * self.avg.get(k, 0.0) + (1.0 - self.decay) * float(v)
        return self.avg


def get_dataset():
    """train.py:112-115."""
    train_imdb = imdb.get_imdb(cfg.train.imdb, is_training=True)
    need_imfeats = cfg.gnet.imfeats or cfg.gnet.load_imfeats
    return ShuffledDataset(train_imdb, 1, need_imfeats), train_imdb


def train(resume, visualize, images_per_step=1):
    if visualize:
        raise NotImplementedError('-v needs matplotlib and the images; not part of this build')
    rank, world, _ = parallel.init_from_env()
    chatty = rank == 0
    np.random.seed(cfg.random_seed)          # every rank draws the same permutation
    dataset, train_imdb = get_dataset()
    do_val = len(cfg.train.val_imdb) > 0

    class_weights = class_equal_weights(train_imdb)
    net = Gnet(num_classes=train_imdb['num_classes'], weight_reg=cfg.train.weight_decay,
               class_weights=class_weights)
    trainer = Trainer(net)
    lr_gen = LearningRate()
    val_net = val_imdb = None
    if do_val:
        val_imdb = imdb.get_imdb(cfg.train.val_imdb, is_training=False)
        val_net = Gnet(num_classes=val_imdb['num_classes'], reuse=True)

    saver = Saver('./')
    model_manager = ModelManager()
    smoothed = SmoothedLosses()
    start_iter = 1
    if resume:
        start_iter = saver.restore_latest(trainer)
        for it in range(1, start_iter):      # replay the stateful schedule
            lr_gen.get_lr(it)
        if chatty:
            print('resuming at iteration {}'.format(start_iter))

    prefetch = Prefetcher(dataset, cfg.train.num_iter - start_iter + 1,
                          images_per_step=images_per_step).start()
    try:
        for it in range(start_iter, cfg.train.num_iter + 1):
            lr = lr_gen.get_lr(it)
            batches = prefetch.get()
            res = trainer.step(parallel.shard(batches), lr)
            # the reference updates its ExponentialMovingAverage on EVERY train step
            # (train.py:263-272, 316-320), not only when it prints
            lo = res['loss_out'].sum(dim=0).cpu().numpy() / max(1, res['num_images']) \
                if res.get('loss_out') is not None else np.zeros(3)
            avg = smoothed.update(total=lo[2] + trainer.regularization_loss(), normed=lo[1],
                                  unnormed=lo[0])
            if it % cfg.train.display_iter == 0 and chatty:
                print(('{}  iter {:6d}   lr {:8g}   opt loss {:8g}     '
                       'data loss normalized {:8g}   unnormalized {:8g}').format(
                    datetime.now(), it, lr, avg['total'], avg['normed'], avg['unnormed']))

            if do_val and it % cfg.train.val_iter == 0:
                if chatty:
                    print('{}  starting validation'.format(datetime.now()))
                    val_map, mc_ap, _ = evaluation.val_run(val_net, val_imdb)
                    print(('{}  iter {:6d}   validation pass:   mAP {:5.1f}   '
                           'multiclass AP {:5.1f}').format(datetime.now(), it, val_map, mc_ap))
                    save_path = saver.save(trainer, net.name, global_step=it)
                    print('wrote model to {}'.format(save_path))
                    model_manager.add(it, val_map, save_path)
                    model_manager.print_summary()
                    model_manager.write_link_to_best('./gnet_best')
            elif it % cfg.train.save_iter == 0 or it == cfg.train.num_iter:
                if chatty:
                    save_path = saver.save(trainer, net.name, global_step=it)
                    print('wrote model to {}'.format(save_path))
            if (do_val and it % cfg.train.val_iter == 0) or it % cfg.train.save_iter == 0:
                # rank 0 validated / saved alone: the others wait HERE, not inside the next
                # step's all-reduce (a long validation would run into the NCCL watchdog)
                parallel.barrier()
    finally:
        prefetch.stop()
    if chatty:
        print('training finished')
        if do_val and model_manager.models:
            print('summary of validation performance')
            model_manager.print_summary()
    return trainer


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument('-r', '--resume', default=False, action='store_true')
    parser.add_argument('-c', '--config', default='conf.yaml')
    parser.add_argument('-v', '--visualize', default=False, action='store_true')
    parser.add_argument('--images-per-step', type=int, default=1)
    args, unparsed = parser.parse_known_args()

    cfg_from_file(args.config)
    if int(os.environ.get('RANK', '0')) == 0:
        pprint(cfg)
    train(args.resume, args.visualize, args.images_per_step)


if __name__ == '__main__':
    main()
