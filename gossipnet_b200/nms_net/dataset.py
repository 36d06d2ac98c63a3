"""roidb -> batches (reference nms_net/dataset.py:17-139).

`load_roi` returns a shallow copy of the roidb entry with `im_scale`; image
loading (`need_images`, ResNet features) is outside the hot path, so boxes keep
scale 1.0 and asking for images raises.  `ShuffledDataset` / `TestDataset`
keep the reference's one-image `next_batch()`; `next_batches(k)` hands the
B200 path k images at once (the per-image sess.run loop of the reference
becomes one batched call).  `Prefetcher` is the reference's load-and-enqueue
thread (train.py:80-109) on a plain queue.
"""
import queue
import threading

import numpy as np

from gossipnet_b200.nms_net.config import cfg  # noqa: F401


def load_roi(need_images, roi, is_training=False):
    """dataset.py:17-45.  With cfg.gnet.imfeats the reference loads the image and runs
    ResNet-101 on it; here the stride-16 feature map travels with the roidb entry as
    `imfeats` [1,H,W,C] (network.py: Gnet._pack_imfeats) and is passed through.  Only an
    entry that would need a real image decoded - no `imfeats` on it - is refused."""
    if need_images and roi.get('imfeats') is None:
        raise NotImplementedError('image loading (load_imfeats, or cfg.gnet.imfeats without a '
                                  'precomputed `imfeats` feature map on the roidb entry) is '
                                  'outside the B200 hot path')
    roi = dict(roi)
    roi['im_scale'] = 1.0
    return roi


class TestDataset(object):
    """dataset.py:70-86: roidb order, one image per batch."""
    __test__ = False   # not a pytest class

    def __init__(self, imdb, batch_size, need_images):
        assert batch_size == 1
        self._imdb, self._roidb = imdb, imdb['roidb']
        self._need_images = need_images
        self._cur = 0

    def next_batch(self):
        i = self._cur
        self._cur += 1
        return load_roi(self._need_images, self._roidb[i], is_training=False)

    def __len__(self):
        return len(self._roidb)


class ShuffledDataset(object):
    """dataset.py:89-112: a fresh np.random permutation per epoch; an epoch's
    tail shorter than the batch size is dropped by reshuffling."""

    def __init__(self, imdb, batch_size, need_images):
        self._imdb, self._roidb = imdb, imdb['roidb']
        self._batch_size = batch_size
        self._need_images = need_images
        self._shuffle()

    def _shuffle(self):
        self._perm = np.random.permutation(np.arange(len(self._roidb)))
        self._cur = 0

    def _take(self, k):
        if self._cur + k > self._perm.size:
            self._shuffle()
        inds = self._perm[self._cur:self._cur + k]
        self._cur += k
        return inds

    def next_batch(self):
        inds = self._take(self._batch_size)
        assert len(inds) == 1
        return load_roi(self._need_images, self._roidb[inds[0]], is_training=True)

    def next_batches(self, k):
        """k images of the current permutation (the batched training step)."""
        return [load_roi(self._need_images, self._roidb[i], is_training=True)
                for i in self._take(k)]


class _LoaderError(object):
    def __init__(self, error):
        self.error = error


class Prefetcher(object):
    """Background thread filling a bounded queue with batches
    (train.py:80-109 / dataset.py:115-139 without the TF queue ops)."""

    def __init__(self, dataset, num_iter, q_size=None, images_per_step=1):
        self.q = queue.Queue(maxsize=q_size or cfg.prefetch_q_size)
        self._stop = threading.Event()
        self._dataset, self._num_iter, self._k = dataset, num_iter, images_per_step
        self._thread = threading.Thread(target=self._run, daemon=True)

    def _put(self, item):
        while not self._stop.is_set():
            try:
                self.q.put(item, timeout=0.1)
                return
            except queue.Full:
                continue

    def _run(self):
        try:
            for _ in range(self._num_iter):
                if self._stop.is_set():
                    return
                self._put(self._dataset.next_batches(self._k) if self._k > 1
                          else [self._dataset.next_batch()])
        except BaseException as e:      # hand the failure to the consumer instead of dying silently
            self._put(_LoaderError(e))

    def start(self):
        self._thread.start()
        return self

    def get(self):
        """Next batch; re-raises an exception of the loader thread, and never blocks forever on
        a thread that is gone."""
        while True:
            try:
                item = self.q.get(timeout=1.0)
            except queue.Empty:
                if not self._thread.is_alive() and self.q.empty():
                    raise RuntimeError('prefetch thread ended without delivering a batch')
                continue
            if isinstance(item, _LoaderError):
                raise item.error
            return item

    def size(self):
        return self.q.qsize()

    def stop(self):
        self._stop.set()
        self._thread.join(timeout=5)
