"""ORACLE (test infrastructure, NOT product code): gradients of the reference's
training objective, by manual backpropagation in float64 numpy.

Follows what TF autodiff generates for nms_net/network.py (no custom gradients
on this path): MatMul/BiasAdd/Relu grads for every fully_connected
(:257-273, :324-342, :344-409), tf.gather grad = scatter-add (:368-369),
tf.select zeroing self-pair rows (:372-374), tf.segment_max grad = rows equal
to the max share the gradient evenly (:387-388), stop_gradient on the geometry
features (:454), DetectionMatching NotDifferentiable
(matching_module/__init__.py:11) -> labels/weights are constants, loss =
sum(weights * sigmoid_ce) (* 1/N if normalize_loss) * loss_multiplyer
(:301-313), L2 regulariser weight_decay * sum(W^2)/2 on FCs built with
weight_reg (train.py:231; not the predict head, network.py:261-272).

tests/test_oracle_grad.py validates this file against central finite
differences of the float64 forward below.
"""
import numpy as np

F64 = np.float64


def _fc(x, p, scope, relu):
    y = x @ p[scope + '/weights'] + p[scope + '/biases']
    return np.maximum(y, 0.0) if relu else y


def forward(params, cfg, pairs, pw_raw, n_dets, keep=False, roi_x0=None):
    """float64 forward from the (constant) raw pair features to the logits.
    `roi_x0` [n_dets, ph*pw*C]: flattened ROI-pooled image features (constants of the
    step) for cfg.gnet.imfeats - the start features are then their reduce_imfeats FCs
    (network.py:223-240) instead of zeros."""
    g = cfg.gnet
    p = dict((k, np.asarray(v, dtype=F64)) for k, v in params.items())
    pc, pn = pairs[:, 0], pairs[:, 1]
    acts = {'pw': [np.asarray(pw_raw, dtype=F64)]}
    for i in range(1, g.num_pwfeat_fc + 1):
        acts['pw'].append(_fc(acts['pw'][-1], p, 'gnet/pw_feats/fc%d' % i, True))
    pw = acts['pw'][-1]
    feats = np.zeros((n_dets, g.shortcut_dim), dtype=F64)
    acts['im'], acts['im_scopes'] = None, []
    if roi_x0 is not None:
        scope = 'gnet/reduce_imfeats/fully_connected'
        acts['im'] = [np.asarray(roi_x0, dtype=F64)]
        if g.imfeat_dim > 0:
            acts['im'].append(_fc(acts['im'][-1], p, scope, True))
            acts['im_scopes'].append(scope)
            scope += '_1'
        acts['im'].append(_fc(acts['im'][-1], p, scope, True))
        acts['im_scopes'].append(scope)
        feats = acts['im'][-1]
    starts = np.flatnonzero(np.diff(np.concatenate([[-1], pc])) != 0)
    blocks = []
    for b in range(1, g.num_blocks + 1):
        s = 'gnet/block%d/' % b
        red = _fc(feats, p, s + 'reduce_dim', True)
        nred = _fc(feats, p, s + 'reduce_dim_neighbor', True) if g.neighbor_feats else red
        nf = nred[pn].copy()
        nf[pc == pn] = 0.0
        x = np.concatenate([pw, red[pc], nf], axis=1)
        hs = [x]
        for i in range(1, g.num_block_pw_fc + 1):
            hs.append(_fc(hs[-1], p, s + 'pw_fc%d' % i, True))
        pooled = np.maximum.reduceat(hs[-1], starts, axis=0)
        ds = [pooled]
        for i in range(1, g.num_block_fc):
            ds.append(_fc(ds[-1], p, s + 'fc%d' % i, True))
        out = np.maximum(feats + _fc(ds[-1], p, s + 'fc%d' % g.num_block_fc, False), 0.0)
        blocks.append((feats, red, nred, hs, ds, out))
        feats = out
    pa = [feats]
    for i in range(1, g.num_predict_fc):
        pa.append(_fc(pa[-1], p, 'gnet/predict/fc%d/fully_connected' % i, False))
    pred = _fc(pa[-1], p, 'gnet/predict/logits/fully_connected', False).reshape(-1)
    if keep:
        return pred, (p, acts, blocks, pa, starts)
    return pred


def data_loss(pred, labels, weights, cfg):
    x, z, w = pred.astype(F64), labels.astype(F64), weights.astype(F64)
    per = (np.maximum(x, 0) - x * z + np.log1p(np.exp(-np.abs(x)))) * w
    base = per.mean() if cfg.train.normalize_loss else per.sum()
    return base * cfg.train.loss_multiplyer


def reg_loss(params, layout, weight_decay):
    return sum(weight_decay * 0.5 * float(np.sum(np.asarray(params[e.name], dtype=F64) ** 2))
               for e in layout.values() if e.regularized)


def gradients(params, cfg, pairs, pw_raw, n_dets, labels, weights, roi_x0=None):
    """d data_loss / d theta for every parameter (float64 dict), plus the logits."""
    g = cfg.gnet
    pred, (p, acts, blocks, pa, starts) = forward(params, cfg, pairs, pw_raw, n_dets, keep=True,
                                                  roi_x0=roi_x0)
    pc, pn = pairs[:, 0], pairs[:, 1]
    grads = dict((k, np.zeros_like(v)) for k, v in p.items())

    def fc_bwd(x, dy, scope):
        grads[scope + '/weights'] += x.T @ dy
        grads[scope + '/biases'] += dy.sum(axis=0)
        return dy @ p[scope + '/weights'].T

    x = pred
    sig = 1.0 / (1.0 + np.exp(-x))
    scale = cfg.train.loss_multiplyer * (1.0 / max(n_dets, 1) if cfg.train.normalize_loss else 1.0)
    d = (scale * weights.astype(F64) * (sig - labels.astype(F64))).reshape(-1, 1)
    d = fc_bwd(pa[-1], d, 'gnet/predict/logits/fully_connected')
    for i in range(g.num_predict_fc - 1, 0, -1):
        d = fc_bwd(pa[i - 1], d, 'gnet/predict/fc%d/fully_connected' % i)
    dfeats = d
    w_pw = acts['pw'][-1].shape[1]
    r = g.reduced_dim
    dpw = np.zeros_like(acts['pw'][-1])
    seg = np.repeat(np.arange(len(starts)), np.diff(np.concatenate([starts, [len(pc)]])))
    for b in range(g.num_blocks, 0, -1):
        s = 'gnet/block%d/' % b
        feats_in, red, nred, hs, ds, out = blocks[b - 1]
        dpre = dfeats * (out > 0)
        dd = fc_bwd(ds[-1], dpre, s + 'fc%d' % g.num_block_fc)
        for i in range(g.num_block_fc - 1, 0, -1):
            dd = dd * (ds[i] > 0)
            dd = fc_bwd(ds[i - 1], dd, s + 'fc%d' % i)
        h = hs[-1]
        sel = (h == ds[0][seg])
        cnt = np.add.reduceat(sel.astype(F64), starts, axis=0)
        dh = np.where(sel, (dd / np.maximum(cnt, 1.0))[seg], 0.0)
        for i in range(g.num_block_pw_fc, 0, -1):
            dh = dh * (hs[i] > 0)
            dh = fc_bwd(hs[i - 1], dh, s + 'pw_fc%d' % i)
        dpw += dh[:, :w_pw]
        dred = np.zeros_like(red)
        np.add.at(dred, pc, dh[:, w_pw:w_pw + r])
        dn = dh[:, w_pw + r:].copy()
        dn[pc == pn] = 0.0
        dnred = np.zeros_like(red) if g.neighbor_feats else dred
        np.add.at(dnred, pn, dn)
        dfeats = dpre + fc_bwd(feats_in, dred * (red > 0), s + 'reduce_dim')
        if g.neighbor_feats:
            dfeats = dfeats + fc_bwd(feats_in, dnred * (nred > 0), s + 'reduce_dim_neighbor')
    if acts['im'] is not None:      # dfeats = d loss / d start features
        d = dfeats
        for i in range(len(acts['im_scopes']), 0, -1):
            d = d * (acts['im'][i] > 0)
            d = fc_bwd(acts['im'][i - 1], d, acts['im_scopes'][i - 1])
    d = dpw
    for i in range(g.num_pwfeat_fc, 0, -1):
        d = d * (acts['pw'][i] > 0)
        d = fc_bwd(acts['pw'][i - 1], d, 'gnet/pw_feats/fc%d' % i)
    return grads, pred


def adam_reference(theta, grad, m, v, lr, step, beta1=0.9, beta2=0.999, eps=1e-8):
    """tf.train.AdamOptimizer update (float64)."""
    lr_t = lr * np.sqrt(1 - beta2 ** step) / (1 - beta1 ** step)
    m = beta1 * m + (1 - beta1) * grad
    v = beta2 * v + (1 - beta2) * grad * grad
    return theta - lr_t * m / (np.sqrt(v) + eps), m, v
