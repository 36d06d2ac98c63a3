"""`Gnet`: the reference's model surface (nms_net/network.py:121-511) on the
B200-native hot path.

Same constructor, static helpers and attribute names as the reference class, so
the consumers in train.py / test.py keep reading `net.prediction`,
`net.labels`, `net.weights`, `net.loss`, `net.det_gt_matching`, ... The
reference builds a TF-0.12 graph once and feeds it through a Session; here the
"graph" is the fixed sequence of C-ABI calls in `gossipnet_b200.engine` and a
"session run" is `net(batch)` (or `net.run(batch)`), which executes it on
`cuda:0` and fills the attributes.  Nothing is computed in Python, and nothing
runs without the CUDA extension.

Differences that are deliberate:
  * more than one image per call: `net.run_batch([batch, ...])` (the reference
    is hard-wired to one image per step, dataset.py:70,105);
  * `det_det_iou` (dense N x N) and `neighbor_pair_idxs` ([P,2] int64) are
    materialised only when read: the hot path keeps the neighbor graph as a
    CSR + int32 pair list and never writes the dense matrix;
  * image features (`cfg.gnet.imfeats`): the head of network.py:223-240 (enlarged
    boxes -> roi_pool -> flatten -> FC -> FC as block-0 features) is here; the
    ResNet-101 that produces the feature map is not, so the
    batch carries the stride-16 feature map itself as `imfeats` [1,H,W,C] where
    the reference's carries the `image`.
"""
import numpy as np
import torch

from gossipnet_b200 import ops
from gossipnet_b200.engine import CapacityOverflow, GnetEngine
from gossipnet_b200.nms_net import matching_module  # noqa: F401  (same import as the reference)
from gossipnet_b200.nms_net.config import cfg

# the reference's batch spec uses tf dtypes; these are the torch equivalents
float32, int32, bool_ = torch.float32, torch.int32, torch.bool

# tf.variable_scope(reuse=True) equivalent: name -> engine of the first Gnet
_SCOPES = {}


class Variable(object):
    """Stand-in for a tf.Variable: `.name` as TF prints it, `.value` a view
    into the engine's flat parameter buffer ([in, out] for weights)."""

    def __init__(self, name, value, regularized):
        self.name = name + ':0'
        self.op_name = name
        self.value = value
        self.regularized = regularized

    def __repr__(self):
        return 'Variable(%s, shape=%s)' % (self.name, tuple(self.value.shape))


def _as_dev(x, dtype, device):
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=dtype).contiguous()
    return torch.from_numpy(np.ascontiguousarray(x)).to(device=device, dtype=dtype).contiguous()


class Gnet(object):
    name = 'gnet'
    dets = None
    det_scores = None
    det_classes = None
    gt_boxes = None
    gt_crowd = None
    gt_classes = None
    image = None

    @staticmethod
    def get_batch_spec(num_classes, is_training=True):
        """network.py:131-146."""
        batch_spec = {
            'dets': (float32, [None, 4]),
            'det_scores': (float32, [None]),
            'det_classes': (int32, [None]),
        }
        if is_training:
            batch_spec.update({
                'gt_boxes': (float32, [None, 4]),
                'gt_crowd': (bool_, [None]),
                'gt_classes': (int32, [None]),
            })
        if cfg.gnet.imfeats:
            # the reference feeds `image` through ResNet-101 (network.py:52-75); here the
            # batch carries that network's stride-16 output map
            batch_spec['imfeats'] = (float32, [1, None, None, cfg.gnet.imfeat_channels])
        elif cfg.gnet.load_imfeats:
            batch_spec['image'] = (float32, [None, None, None, 3])
        return batch_spec

    def __init__(self, num_classes, class_weights=None, batch=None,
                 weight_reg=None, reuse=False, device='cuda', params=None):
        self.num_classes = num_classes
        self.multiclass = num_classes > 1
        self.weight_reg = weight_reg
        if reuse:
            if self.name not in _SCOPES:
                raise ValueError('Gnet(reuse=True) before any Gnet was built '
                                 '(variable scope gnet does not exist)')
            self.engine = _SCOPES[self.name]
            if self.engine.num_classes != num_classes:
                raise ValueError('Gnet(reuse=True) with a different num_classes')
        else:
            self.engine = GnetEngine(num_classes, cfg, device=device, flat_params=params)
            _SCOPES[self.name] = self.engine
        self.device = self.engine.device
        if class_weights is None:
            class_weights = np.ones((num_classes + 1), dtype=np.float32)  # network.py:282-283
        self.class_weights = _as_dev(np.asarray(class_weights, dtype=np.float32), float32,
                                     self.device)
        # network.py:316-322
        self.trainable_variables = [
            Variable(e.name, self.engine.p[e.name], e.regularized)
            for e in self.engine.layout.values()]
        self._clear()
        if batch is not None:
            self.run(batch)

    # ------------------------------------------------------------------ plumbing
    def _clear(self):
        for k in ('prediction', 'labels', 'weights', 'det_gt_matching', 'loss', 'loss_normed',
                  'loss_unnormed', 'det_anno_iou', 'pw_feats', 'block_feats', 'num_dets',
                  'dets_boxdata', 'gt_boxdata', 'imfeats', 'roifeats', 'det_imfeats',
                  'frcn_boxes'):
            setattr(self, k, None)
        self._lazy = {}

    def state_dict(self):
        """name -> [in,out] / [out] tensors (views of the flat buffer), keyed by
        the reference's TF variable names (`gnet/block3/pw_fc1/weights`, ...)."""
        return dict((v.op_name, v.value) for v in self.trainable_variables)

    def load_state_dict(self, sd):
        for v in self.trainable_variables:
            v.value.copy_(_as_dev(sd[v.op_name], float32, self.device))

    @staticmethod
    def _pack(batches, device, with_gt):
        """Concatenate per-image host/device arrays -> device tensors + offsets."""
        n = [int(np.shape(b['dets'])[0]) for b in batches]
        img_off_host = np.zeros(len(batches) + 1, dtype=np.int32)
        np.cumsum(n, out=img_off_host[1:])
        cat = lambda key, dt: torch.cat([_as_dev(b[key], dt, device).reshape(
            (-1, 4) if key in ('dets', 'gt_boxes') else (-1,)) for b in batches])
        out = dict(dets=cat('dets', float32), det_scores=cat('det_scores', float32),
                   det_classes=cat('det_classes', int32), img_off_host=img_off_host,
                   img_off=torch.from_numpy(img_off_host).to(device))
        if with_gt:
            g = [int(np.shape(b['gt_boxes'])[0]) if np.size(b['gt_boxes']) else 0 for b in batches]
            gt_off_host = np.zeros(len(batches) + 1, dtype=np.int32)
            np.cumsum(g, out=gt_off_host[1:])
            out.update(gt_boxes=cat('gt_boxes', float32),
                       gt_crowd=torch.cat([_as_dev(np.asarray(b['gt_crowd']).astype(np.uint8)
                                                   if not isinstance(b['gt_crowd'], torch.Tensor)
                                                   else b['gt_crowd'].to(torch.uint8),
                                                   torch.uint8, device).reshape(-1)
                                           for b in batches]),
                       gt_classes=cat('gt_classes', int32), gt_off_host=gt_off_host)
        return out

    def _pack_imfeats(self, batches):
        """Per-image feature maps [1,H,W,C] on the device (None without cfg.gnet.imfeats)."""
        if not self.engine.imfeats:
            return None
        if any(b.get('imfeats') is None for b in batches):
            raise NotImplementedError(
                'cfg.gnet.imfeats: every image needs its stride-16 feature map as '
                '`imfeats` [1,H,W,C]; computing it from `image` (ResNet-101) is not part '
                'of this package')
        return [_as_dev(b['imfeats'], float32, self.device) for b in batches]

    # ----------------------------------------------------------------------- run
    def run_batch(self, batches, want_grad=False):
        """Forward (+ matching and loss when every image carries gt_*) for a list
        of images.  Returns a dict of device tensors over the concatenated
        detections; `img_off_host` gives each image's rows."""
        with_gt = all(b.get('gt_boxes') is not None for b in batches)
        io = self._pack(batches, self.device, with_gt)
        eng = self.engine
        imfeats = io['imfeats'] = self._pack_imfeats(batches)
        while True:
            res = eng.forward(io['dets'], io['det_scores'], io['det_classes'], io['img_off'],
                              imfeats=imfeats, img_off_host=io['img_off_host'])
            if with_gt and io['dets'].shape[0] > 0:
                res.update(eng.matching_and_loss(
                    res['prediction'], io['dets'], io['det_classes'], io['img_off'],
                    io['img_off_host'], io['gt_boxes'], io['gt_crowd'], io['gt_classes'],
                    io['gt_off_host'], self.class_weights, want_grad=want_grad))
            try:
                res['P'] = eng.check_overflow()
                break
            except CapacityOverflow:
                continue
        res.update(io)
        return res

    def run(self, batch):
        """One image, like one `sess.run` of the reference graph: fills the
        attributes and returns `prediction`."""
        self._clear()
        eng = self.engine
        eng.keep_block_feats = True
        try:
            res = self.run_batch([batch])
        finally:
            eng.keep_block_feats = False
        self._res = res
        for k in ('dets', 'det_scores', 'det_classes'):
            setattr(self, k, res[k])
        self.num_dets = res['dets'].shape[0]
        P = res['P']
        self.prediction = res['prediction'].clone()
        self.pw_feats = res['pw_feats'][:P].clone()
        self.block_feats = res['block_feats']
        self._pairs = (res['pair_c'][:P].clone(), res['pair_n'][:P].clone())
        self.dets_boxdata = self._xyxy_to_boxdata(self.dets)
        if 'roifeats' in res:
            self.imfeats = res['imfeats'][0]
            self.roifeats, self.frcn_boxes = res['roifeats'], res['frcn_boxes'].clone()
            self.det_imfeats = res['det_imfeats'].clone()
        if 'labels' in res:
            for k in ('gt_boxes', 'gt_crowd', 'gt_classes'):
                setattr(self, k, res[k])
            self.gt_boxdata = self._xyxy_to_boxdata(self.gt_boxes)
            self.labels, self.weights = res['labels'], res['weights']
            self.det_gt_matching = res['det_gt_matching']
            self.det_anno_iou = res['det_anno_iou'].reshape(
                self.num_dets, int(res['gt_boxes'].shape[0])).clone()
            lo = res['loss_out'][0]
            self.loss_unnormed, self.loss_normed, self.loss = lo[0], lo[1], lo[2]
        return self.prediction

    __call__ = run

    # ------------------------------------------------------- lazily built tensors
    @property
    def neighbor_pair_idxs(self):
        """[P,2] int64, row-major order, as tf.where gives it (network.py:192-195)."""
        if 'pairs' not in self._lazy:
            c, n = self._pairs
            self._lazy['pairs'] = torch.stack([c, n], dim=1).to(torch.int64)
        return self._lazy['pairs']

    @property
    def det_det_iou(self):
        """Dense [N,N] overlap matrix (network.py:176); off the hot path."""
        if 'ddi' not in self._lazy:
            self._lazy['ddi'] = ops.iou_dense(self.dets, self.dets)
        return self._lazy['ddi']

    # ------------------------------------------------- the reference's static API
    @staticmethod
    def _xyxy_to_boxdata(a):
        """network.py:462-472 (attribute parity only; kernels recompute it)."""
        x1, y1, x2, y2 = a[:, 0:1], a[:, 1:2], a[:, 2:3], a[:, 3:4]
        w, h = x2 - x1, y2 - y1
        return (x1, y1, w, h, x2, y2, w * h)

    @staticmethod
    def _iou(a, b, crowd=None):
        """network.py:474-488 on boxdata tuples -> dense [n,m] (CUDA)."""
        box = lambda t: torch.cat([t[0], t[1], t[4], t[5]], dim=1).contiguous()
        return ops.iou_dense(box(a), box(b), crowd=crowd)

    @staticmethod
    def _block(block_idx, infeats, weights_init, biases_init,
               pair_c_idxs, pair_n_idxs, pw_feats, weight_reg):
        """network.py:344-409 as a standalone call on the current `gnet` scope's
        parameters.  `weights_init` / `biases_init` / `weight_reg` are accepted
        for signature parity (parameters already exist here)."""
        if Gnet.name not in _SCOPES:
            raise ValueError('Gnet._block needs a constructed Gnet (its variables)')
        eng = _SCOPES[Gnet.name]
        dev = eng.device
        infeats = _as_dev(infeats, float32, dev)
        pc = _as_dev(pair_c_idxs, int32, dev)
        pn = _as_dev(pair_n_idxs, int32, dev)
        pw = _as_dev(pw_feats, float32, dev)
        T, npairs = infeats.shape[0], pc.numel()
        num_pairs = torch.tensor([npairs], dtype=torch.int32, device=dev)
        # CSR row pointer from the sorted c indices (segment ids)
        counts = torch.bincount(pc.to(torch.int64), minlength=T).to(torch.int32)
        row_ptr = ops.exclusive_scan(counts)
        out = torch.empty_like(infeats)
        return eng.block(block_idx, infeats, row_ptr, pc, pn, num_pairs, npairs, pw, out)
