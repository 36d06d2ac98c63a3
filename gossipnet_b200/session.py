"""Host-buffer inference sessions: the batched equivalent of the reference's
test loop (test.py:63-71: feed numpy arrays -> sess.run(net.prediction) -> numpy
scores), with the host<->device copies done from pinned staging buffers and the
whole forward replayed as ONE CUDA graph per batch shape.

    sess = InferenceSession(net)
    new_scores = sess.run(dets, det_scores, det_classes, img_off)   # numpy in, numpy out

`run` is the call bench.py times for the end-to-end number: it includes the
host->device copy of the step's inputs and the device->host read of its
logits.
"""
import numpy as np
import torch

from gossipnet_b200 import _lib
from gossipnet_b200.engine import CapacityOverflow


class InferenceSession(object):

    def __init__(self, net, use_graph=True):
        self.net = net
        self.engine = net.engine
        self.device = net.device
        self.use_graph = use_graph
        self._shape = None
        self._max_img = 0
        self._graph = None
        self._graph_io = None
        self.stream = torch.cuda.Stream(device=self.device)
        self.launches_per_forward = None

    # ------------------------------------------------------------------ buffers
    def _alloc(self, T, B):
        """One pinned staging buffer and one device buffer hold all four inputs (typed views at
        256-byte aligned offsets), so a step's inputs travel in ONE host->device copy."""
        dev = self.device
        sizes = [('dets', T * 16), ('scores', T * 4), ('cls', T * 4), ('off', (B + 1) * 4)]
        offs, total = {}, 0
        for name, nbytes in sizes:
            offs[name] = total
            total += (nbytes + 255) // 256 * 256
        self.h_in = torch.empty(max(total, 256), dtype=torch.uint8, pin_memory=True)
        self.d_in = torch.empty(max(total, 256), dtype=torch.uint8, device=dev)

        def view(buf, name, nbytes, dtype, shape):
            return buf[offs[name]:offs[name] + nbytes].view(dtype).view(*shape)
        for buf, pre in ((self.h_in, 'h_'), (self.d_in, 'd_')):
            setattr(self, pre + 'dets', view(buf, 'dets', T * 16, torch.float32, (T, 4)))
            setattr(self, pre + 'scores', view(buf, 'scores', T * 4, torch.float32, (T,)))
            setattr(self, pre + 'cls', view(buf, 'cls', T * 4, torch.int32, (T,)))
            setattr(self, pre + 'off', view(buf, 'off', (B + 1) * 4, torch.int32, (B + 1,)))
        self.h_pred = torch.empty((T,), dtype=torch.float32, pin_memory=True)
        self.h_np = torch.empty((1,), dtype=torch.int32, pin_memory=True)
        self.h2d_bytes = T * (16 + 4 + 4) + (B + 1) * 4
        self.d2h_bytes = T * 4 + 4
        self._shape = (T, B)
        self._graph = None

    def _forward(self):
        res = self.engine.forward(self.d_dets, self.d_scores, self.d_cls, self.d_off,
                                  max_img=self._max_img)
        self._pred, self._num_pairs, self._cap = res['prediction'], res['num_pairs'], res['capacity']

    def _prepare(self):
        """Warm up (learns the pair capacity, sets kernel attributes) and capture."""
        with torch.cuda.stream(self.stream):
            for _ in range(2):
                while True:
                    self._forward()
                    try:
                        self.engine.check_overflow()
                        break
                    except CapacityOverflow:
                        continue
            n0 = _lib.CALLS[0]
            self._forward()
            self.launches_per_forward = _lib.CALLS[0] - n0
            self.stream.synchronize()
            if self.use_graph:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=self.stream):
                    self._forward()
                self._graph = g
                # the whole host-buffer step as one graph: H2D of the staging buffer, the
                # forward, D2H of the logits and of the pair count (one launch per run())
                gio = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gio, stream=self.stream):
                    self.d_in.copy_(self.h_in, non_blocking=True)
                    self._forward()
                    self.h_pred.copy_(self._pred, non_blocking=True)
                    self.h_np.copy_(self._num_pairs, non_blocking=True)
                self._graph_io = gio
            else:
                self._graph = False
                self._graph_io = None

    # ---------------------------------------------------------------------- run
    def run(self, dets, det_scores, det_classes, img_off):
        """numpy: dets[T,4] f32, det_scores[T] f32, det_classes[T] i32,
        img_off[B+1] i32 -> new scores (logits) [T] f32 (a view of the pinned
        result buffer, valid until the next run)."""
        T, B = int(dets.shape[0]), int(img_off.shape[0]) - 1
        # the captured graph is specific to (T, B) and to the mask stride of the neighbor
        # build, i.e. to the size of the largest image (rounded up to 32 detections)
        max_img = (int(np.max(np.diff(img_off))) + 31) // 32 * 32 if B > 0 else 0
        if self._shape != (T, B) or max_img != self._max_img:
            self._max_img = max_img
            self._alloc(T, B)
        self.h_dets.numpy()[...] = dets
        self.h_scores.numpy()[...] = det_scores
        self.h_cls.numpy()[...] = det_classes
        self.h_off.numpy()[...] = img_off
        for attempt in range(3):
            with torch.cuda.stream(self.stream):
                if self._graph is None:
                    self.d_in.copy_(self.h_in, non_blocking=True)   # warm-up runs on real inputs
                    self._prepare()
                if self._graph:
                    self._graph_io.replay()
                else:
                    self.d_in.copy_(self.h_in, non_blocking=True)
                    self._forward()
                    self.h_pred.copy_(self._pred, non_blocking=True)
                    self.h_np.copy_(self._num_pairs, non_blocking=True)
            self.stream.synchronize()
            if int(self.h_np[0]) <= self._cap:
                return self.h_pred.numpy()
            # denser batch than the workspace was sized for: grow, re-capture, redo
            self.engine.capacity = int(int(self.h_np[0]) * 1.25) + 256
            self._graph = None
        raise RuntimeError('pair capacity did not converge')
