"""Parameter inventory of Gnet: names, shapes, flat-buffer layout, init.

Names follow the TF variable scopes the reference creates
(`network.py:217-221` pw_feats/fc{i}; `:344-409` block{b}/{reduce_dim,pw_fc1,
pw_fc2,fc1,fc2}; `:257-273` predict/fc{i}/fully_connected and
predict/logits/fully_connected), weights are `[in, out]`, `y = act(x @ W + b)`.
All parameters live in ONE flat float32 buffer (one NCCL all-reduce and one
fused Adam launch per training step); each entry records its
offset into that buffer.
"""
from collections import OrderedDict

import numpy as np


class ParamEntry(object):
    __slots__ = ('name', 'shape', 'offset', 'size', 'is_weight', 'regularized')

    def __init__(self, name, shape, offset, is_weight, regularized):
        self.name = name
        self.shape = tuple(int(s) for s in shape)
        self.offset = int(offset)
        self.size = int(np.prod(self.shape))
        self.is_weight = bool(is_weight)
        self.regularized = bool(regularized)


def raw_pairfeat_width(num_classes):
    """Width of `_geometry_feats` output (`network.py:411-454`): 9 for a
    single class, 2C+7 when scores are one-hot rows."""
    return 9 if num_classes <= 1 else 2 * num_classes + 7


def param_layout(num_classes, cfg):
    """Ordered name -> ParamEntry.  With cfg.gnet.imfeats the image-feature head
    (`network.py:223-240`, scope gnet/reduce_imfeats: flattened ROI features ->
    [imfeat_dim ->] shortcut_dim) is included; the ResNet that produces the
    feature map is not part of this package."""
    g = cfg.gnet
    entries = OrderedDict()
    off = [0]

    def add_fc(scope, n_in, n_out, regularized):
        for leaf, shape, is_w in (('weights', (n_in, n_out), True),
                                  ('biases', (n_out,), False)):
            name = scope + '/' + leaf
            e = ParamEntry(name, shape, off[0], is_w, regularized and is_w)
            # keep every tensor 16-byte aligned for vector loads
            off[0] += (e.size + 3) // 4 * 4
            entries[name] = e

    width = raw_pairfeat_width(num_classes)
    if g.num_pwfeat_fc > 0:
        for i in range(1, g.num_pwfeat_fc):
            add_fc('gnet/pw_feats/fc%d' % i, width, g.pwfeat_dim, True)
            width = g.pwfeat_dim
        add_fc('gnet/pw_feats/fc%d' % g.num_pwfeat_fc, width,
               g.pwfeat_narrow_dim, True)
        width = g.pwfeat_narrow_dim
    if g.imfeats:
        n_in = cfg.imfeat_crop_height * cfg.imfeat_crop_width * g.imfeat_channels
        scope = 'gnet/reduce_imfeats/fully_connected'
        if g.imfeat_dim > 0:
            add_fc(scope, n_in, g.imfeat_dim, True)
            n_in, scope = g.imfeat_dim, scope + '_1'      # TF's second default scope name
        add_fc(scope, n_in, g.shortcut_dim, True)
    for b in range(1, g.num_blocks + 1):
        s = 'gnet/block%d/' % b
        add_fc(s + 'reduce_dim', g.shortcut_dim, g.reduced_dim, True)
        if g.neighbor_feats:
            add_fc(s + 'reduce_dim_neighbor', g.shortcut_dim, g.reduced_dim, True)
        n_in = width + 2 * g.reduced_dim
        for i in range(1, g.num_block_pw_fc + 1):
            add_fc(s + 'pw_fc%d' % i, n_in, g.pairfeat_dim, True)
            n_in = g.pairfeat_dim
        for i in range(1, g.num_block_fc):
            add_fc(s + 'fc%d' % i, n_in, g.pairfeat_dim, True)
            n_in = g.pairfeat_dim
        add_fc(s + 'fc%d' % g.num_block_fc, n_in, g.shortcut_dim, True)
    n_in = g.shortcut_dim
    for i in range(1, g.num_predict_fc):
        add_fc('gnet/predict/fc%d/fully_connected' % i, n_in, g.predict_fc_dim, False)
        n_in = g.predict_fc_dim
    add_fc('gnet/predict/logits/fully_connected', n_in, 1, False)
    return entries, off[0]


def num_params(layout):
    return sum(e.size for e in layout.values())


def init_flat(layout, total, cfg, seed=None):
    """Seeded xavier-uniform weights (+-sqrt(6/(in+out))) and constant biases
    (`network.py:202-214`); TF's own random stream is not reproducible without
    TF, so parity tests share THIS generator between the oracle and CUDA."""
    g = cfg.gnet
    rs = np.random.RandomState(cfg.random_seed if seed is None else seed)
    flat = np.zeros(total, dtype=np.float32)
    for e in layout.values():
        if e.is_weight:
            n_in, n_out = e.shape
            if g.weight_init == 'xavier':
                lim = np.sqrt(6.0 / (n_in + n_out))
                w = rs.uniform(-lim, lim, e.shape)
            elif g.weight_init == 'caffe':
                lim = np.sqrt(3.0 / n_in)
                w = rs.uniform(-lim, lim, e.shape)
            elif g.weight_init == 'msra':
                w = rs.normal(0.0, np.sqrt(2.0 / n_in), e.shape)
            else:
                raise ValueError('unknown weight init {}'.format(g.weight_init))
            flat[e.offset:e.offset + e.size] = w.astype(np.float32).ravel()
        else:
            flat[e.offset:e.offset + e.size] = np.float32(g.bias_const_init)
    return flat


def views(layout, flat):
    """name -> array view of `flat` (numpy array or torch tensor)."""
    return OrderedDict(
        (e.name, flat[e.offset:e.offset + e.size].reshape(e.shape))
        for e in layout.values())
