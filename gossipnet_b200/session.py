"""Host-buffer inference sessions: the batched equivalent of the reference's
test loop (test.py:63-71: feed numpy arrays -> sess.run(net.prediction) -> numpy
scores), with the host<->device copies done from pinned staging buffers.

    sess = InferenceSession(net)
    new_scores = sess.run(dets, det_scores, det_classes, img_off)   # numpy in, numpy out

A batch shape (detections, images, size of the largest image) that comes back is
replayed as ONE CUDA graph (H2D of the staging buffer, the whole forward, D2H of
the logits): the graph is captured the second time the shape is seen, a few
shapes are kept.  Shapes seen once - real detection files have a different
number of detections in almost every batch - run the same calls eagerly through
grow-only staging buffers, so nothing is captured or allocated per call.

`run` is the call bench.py times for the end-to-end number: it includes the
host->device copy of the step's inputs and the device->host read of its logits.
"""
import collections

import numpy as np
import torch

from gossipnet_b200 import _lib
from gossipnet_b200.engine import CapacityOverflow


class _Buffers(object):
    """Pinned + device staging for up to (T, B): one buffer holds all four inputs (typed
    views at 256-byte aligned offsets), so a step's inputs travel in ONE host->device copy."""

    def __init__(self, T, B, device):
        self.T_cap, self.B_cap = T, B
        sizes = [('dets', T * 16), ('scores', T * 4), ('cls', T * 4), ('off', (B + 1) * 4)]
        self.offs, total = {}, 0
        for name, nbytes in sizes:
            self.offs[name] = total
            total += (nbytes + 255) // 256 * 256
        self.h_in = torch.empty(max(total, 256), dtype=torch.uint8, pin_memory=True)
        self.d_in = torch.empty(max(total, 256), dtype=torch.uint8, device=device)
        self.h_pred_all = torch.empty((max(T, 1),), dtype=torch.float32, pin_memory=True)
        self.h_np = torch.empty((1,), dtype=torch.int32, pin_memory=True)
        # per-slot device copy of the results (run_pipelined: the D2H of batch i runs on its
        # own stream while the next forward already overwrites the engine's logits)
        self.d_pred_all = torch.empty((max(T, 1),), dtype=torch.float32, device=device)
        self.d_np = torch.empty((1,), dtype=torch.int32, device=device)
        self.bind(T, B)

    def bind(self, T, B):
        """Typed views for a batch of T detections in B images (T <= T_cap, B <= B_cap)."""
        def view(buf, name, nbytes, dtype, shape):
            o = self.offs[name]
            return buf[o:o + nbytes].view(dtype).view(*shape)
        for buf, pre in ((self.h_in, 'h_'), (self.d_in, 'd_')):
            setattr(self, pre + 'dets', view(buf, 'dets', T * 16, torch.float32, (T, 4)))
            setattr(self, pre + 'scores', view(buf, 'scores', T * 4, torch.float32, (T,)))
            setattr(self, pre + 'cls', view(buf, 'cls', T * 4, torch.int32, (T,)))
            setattr(self, pre + 'off', view(buf, 'off', (B + 1) * 4, torch.int32, (B + 1,)))
        self.h_pred = self.h_pred_all[:T]
        self.d_pred = self.d_pred_all[:T]
        self.T, self.B = T, B
        # bytes that matter (the aligned staging buffer carries a little padding on top)
        self.h2d_bytes = T * (16 + 4 + 4) + (B + 1) * 4
        self.d2h_bytes = T * 4 + 4
        self.copy_bytes = self.offs['off'] + (B + 1) * 4


class InferenceSession(object):
    MAX_GRAPHS = 4          # captured shapes kept (least recently used goes first)

    def __init__(self, net, use_graph=True, graph_after=2):
        self.net = net
        self.engine = net.engine
        self.device = net.device
        self.use_graph = use_graph
        self.graph_after = graph_after       # capture a shape when it is seen this many times
        self.stream = torch.cuda.Stream(device=self.device)
        # copy streams of run_pipelined: inputs of batch i+1 go up and logits of batch i-1 come
        # down while the forward of batch i runs
        self.h2d_stream = torch.cuda.Stream(device=self.device)
        self.d2h_stream = torch.cuda.Stream(device=self.device)
        self.launches_per_forward = None
        self._seen = collections.Counter()
        self._states = collections.OrderedDict()    # shape key -> (buffers, graph, graph_io)
        self._pipe = {}                               # (shape key, slot) -> (buffers, graph_io, ws generation)
        self._eager = None                            # grow-only buffers of the eager path
        self._cur = None                              # buffers of the last run
        self._graph = None                            # forward-only graph of the last run (or None)
        self._max_img = 0

    # attribute surface bench.py and the tests read -----------------------------------
    def __getattr__(self, name):
        if name in ('d_dets', 'd_scores', 'd_cls', 'd_off', 'h_dets', 'h_scores', 'h_cls', 'h_off',
                    'h_np', 'h_pred', 'h2d_bytes', 'd2h_bytes'):
            cur = self.__dict__.get('_cur')
            if cur is None:
                raise AttributeError('%s: no batch has been run yet' % name)
            return getattr(cur, name)
        raise AttributeError(name)

    # ------------------------------------------------------------------- forward
    def _forward(self, buf):
        eng = self.engine
        keep, eng.want_pw_f32 = eng.want_pw_f32, False    # logits only: no fp32 pw_feats copy
        try:
            res = eng.forward(buf.d_dets, buf.d_scores, buf.d_cls, buf.d_off,
                              max_img=self._max_img)
        finally:
            eng.want_pw_f32 = keep
        self._pred, self._num_pairs, self._cap = res['prediction'], res['num_pairs'], res['capacity']

    def _warm(self, buf):
        """Forward until the pair capacity fits (the engine learns it on its first batches)."""
        while True:
            self._forward(buf)
            try:
                self.engine.check_overflow()
                return
            except CapacityOverflow:
                continue

    def _capture(self, buf, pipelined=False):
        """Warm up on the real inputs (already in buf.d_in), count launches, capture the
        forward alone (bench's resident measurement) and the whole host-buffer step
        (pipelined: the forward + device copies of its results into the slot; the host
        copies run on the copy streams)."""
        for _ in range(2):
            self._warm(buf)
        self.engine.refresh_weight_images()     # fresh now: the captures below skip them
        n0 = _lib.CALLS[0]
        self._forward(buf)
        self.launches_per_forward = _lib.CALLS[0] - n0
        self.stream.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=self.stream):
            self._forward(buf)
        gio = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gio, stream=self.stream):
            if pipelined:
                self._forward(buf)
                buf.d_pred.copy_(self._pred.view(-1), non_blocking=True)
                buf.d_np.copy_(self._num_pairs.view(-1), non_blocking=True)
            else:
                buf.d_in[:buf.copy_bytes].copy_(buf.h_in[:buf.copy_bytes], non_blocking=True)
                self._forward(buf)
                buf.h_pred.copy_(self._pred, non_blocking=True)
                buf.h_np.copy_(self._num_pairs, non_blocking=True)
        return g, gio

    def _eager_buffers(self, T, B):
        e = self._eager
        if e is None or T > e.T_cap or B > e.B_cap:
            grow = lambda need, have: max(need, int(have * 1.5))
            e = self._eager = _Buffers(grow(T, e.T_cap if e else 0), grow(B, e.B_cap if e else 0),
                                       self.device)
        e.bind(T, B)
        return e

    # ---------------------------------------------------------------------- run
    def run(self, dets, det_scores, det_classes, img_off):
        """numpy: dets[T,4] f32, det_scores[T] f32, det_classes[T] i32,
        img_off[B+1] i32 -> new scores (logits) [T] f32 (a view of the pinned
        result buffer, valid until the next run)."""
        T, B = int(dets.shape[0]), int(img_off.shape[0]) - 1
        if T == 0:
            return np.zeros((0,), dtype=np.float32)
        # a graph is specific to (T, B) and to the mask stride of the neighbor build, i.e.
        # to the size of the largest image (rounded up to 32 detections)
        self._max_img = (int(np.max(np.diff(img_off))) + 31) // 32 * 32
        key = (T, B, self._max_img)
        self._seen[key] += 1
        if len(self._seen) > 4096:           # variable-size data: do not grow without bound
            self._seen.clear()
        state = self._states.get(key)
        if state is not None and state[3] != self.engine.ws_generation:
            # the engine reallocated part of its workspace since this graph was captured (a
            # larger batch came through): the graph holds pointers into freed memory - drop it
            del self._states[key]
            state = None
        want_graph = self.use_graph and (state is not None or self._seen[key] >= self.graph_after)
        if want_graph and state is None:
            buf = _Buffers(T, B, self.device)
        elif state is not None:
            buf = state[0]
            self._states.move_to_end(key)
        else:
            buf = self._eager_buffers(T, B)
        buf.h_dets.numpy()[...] = dets
        buf.h_scores.numpy()[...] = det_scores
        buf.h_cls.numpy()[...] = det_classes
        buf.h_off.numpy()[...] = img_off
        self._cur = buf
        for attempt in range(4):
            with torch.cuda.stream(self.stream):
                if want_graph and state is None:
                    buf.d_in[:buf.copy_bytes].copy_(buf.h_in[:buf.copy_bytes], non_blocking=True)
                    state = (buf,) + self._capture(buf) + (self.engine.ws_generation,)
                    self._states[key] = state
                    while len(self._states) > self.MAX_GRAPHS:
                        self._states.popitem(last=False)
                if want_graph:
                    self._graph = state[1]
                    # the captured forward does not contain the weight-derived launches
                    # (operand images, folded predict head): redo them if the weights moved
                    self.engine.refresh_weight_images()
                    state[2].replay()
                else:
                    self._graph = False if not self.use_graph else None
                    buf.d_in[:buf.copy_bytes].copy_(buf.h_in[:buf.copy_bytes], non_blocking=True)
                    if self.launches_per_forward is None:
                        self._warm(buf)
                        n0 = _lib.CALLS[0]
                        self._forward(buf)
                        self.launches_per_forward = _lib.CALLS[0] - n0
                    else:
                        self._forward(buf)
                    buf.h_pred.copy_(self._pred, non_blocking=True)
                    buf.h_np.copy_(self._num_pairs, non_blocking=True)
            self.stream.synchronize()
            if int(buf.h_np[0]) <= self._cap:
                return buf.h_pred.numpy()
            # denser batch than the workspace was sized for: grow, drop the graphs (they were
            # captured with the old capacity), redo
            self.engine.capacity = int(int(buf.h_np[0]) * 1.25) + 256
            self._states.clear()
            self._pipe.clear()
            state = None
        raise RuntimeError('pair capacity did not converge')

    # ------------------------------------------------------------- pipelined run
    def _pipe_state(self, key, slot):
        """Staging buffers + captured forward of one pipeline slot.  The two slots of a shape
        have their own pinned and device staging (inputs and results); both graphs run on the
        session's stream, one after the other, over the engine's single workspace."""
        st = self._pipe.get((key, slot))
        if st is not None and st[2] == self.engine.ws_generation:
            return st
        T, B, _ = key
        buf = _Buffers(T, B, self.device)
        return buf, None, None

    def run_pipelined(self, batches):
        """Generator over batches (dets, det_scores, det_classes, img_off) -> new scores per
        batch, in order.  One batch ahead: while the GPU works on batch i the host stages
        batch i+1 into the other slot's pinned buffer, its inputs go up on the H2D stream and
        the logits of batch i-1 come down on the D2H stream, so the compute stream runs the
        captured forwards back to back and per-batch cost is max(host staging, forward)
        instead of the sum of staging, copies and forward.  Every batch's inputs still travel
        pinned host -> device and its logits device -> pinned host.  A yielded array is a view
        of the slot's pinned result buffer: valid until two more batches have been yielded.
        Shapes seen for the first time (or with use_graph=False) go through `run`."""
        pending = None          # (buf, event, args)
        slot = 0

        def finish(p):
            buf, ev, args = p
            ev.synchronize()
            if int(buf.h_np[0]) > self._cap:
                return None     # denser batch than the workspace was sized for
            return buf.h_pred.numpy()

        it = iter(batches)
        for args in it:
            dets, det_scores, det_classes, img_off = args
            T, B = int(dets.shape[0]), int(img_off.shape[0]) - 1
            key = (T, B, (int(np.max(np.diff(img_off))) + 31) // 32 * 32) if T > 0 else None
            graphable = self.use_graph and T > 0
            st = self._pipe_state(key, slot) if graphable else None
            if st is None or st[1] is None:
                # no captured step for this slot yet: drain the pipeline, then set it up
                if pending is not None:
                    out = finish(pending)
                    yield out if out is not None else self.run(*pending[2])
                    pending = None
                if not graphable:
                    yield self.run(*args)
                    continue
                self._max_img = key[2]
                buf = st[0]
                self._fill(buf, args)
                with torch.cuda.stream(self.stream):
                    buf.d_in[:buf.copy_bytes].copy_(buf.h_in[:buf.copy_bytes], non_blocking=True)
                    _, gio = self._capture(buf, pipelined=True)
                self.stream.synchronize()
                self._pipe[(key, slot)] = (buf, gio, self.engine.ws_generation)
                st = self._pipe[(key, slot)]
            buf, gio, _ = st
            # the slot's previous batch (two back) has been yielded: its copies are complete
            self._fill(buf, args)
            self._cur = buf
            with torch.cuda.stream(self.h2d_stream):
                buf.d_in[:buf.copy_bytes].copy_(buf.h_in[:buf.copy_bytes], non_blocking=True)
                ev_in = torch.cuda.Event()
                ev_in.record(self.h2d_stream)
            with torch.cuda.stream(self.stream):
                self.stream.wait_event(ev_in)
                self.engine.refresh_weight_images()
                gio.replay()
                ev_fwd = torch.cuda.Event()
                ev_fwd.record(self.stream)
            with torch.cuda.stream(self.d2h_stream):
                self.d2h_stream.wait_event(ev_fwd)
                buf.h_pred.copy_(buf.d_pred, non_blocking=True)
                buf.h_np.copy_(buf.d_np, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.d2h_stream)
            if pending is not None:
                out = finish(pending)
                if out is None:
                    # capacity overflow: everything in flight was computed with too small a
                    # workspace - grow, drop the captures, redo both batches synchronously
                    ev.synchronize()
                    self.engine.capacity = int(int(pending[0].h_np[0]) * 1.25) + 256
                    self._states.clear()
                    self._pipe.clear()
                    yield self.run(*pending[2]).copy()
                    yield self.run(*args).copy()
                    pending = None
                    continue
                yield out
            pending = (buf, ev, args)
            slot ^= 1
        if pending is not None:
            out = finish(pending)
            yield out if out is not None else self.run(*pending[2])

    @staticmethod
    def _fill(buf, args):
        dets, det_scores, det_classes, img_off = args
        buf.h_dets.numpy()[...] = dets
        buf.h_scores.numpy()[...] = det_scores
        buf.h_cls.numpy()[...] = det_classes
        buf.h_off.numpy()[...] = img_off
