"""CPU: the gradient oracle (manual backprop, oracle/gnet_grad_oracle.py) against
central finite differences of its own float64 forward, on a tiny graph."""
import numpy as np

from gossipnet_b200 import params as P
from gossipnet_b200 import synthetic
from gossipnet_b200.nms_net.config import cfg
from oracle import gnet_grad_oracle as gg
from oracle import gnet_oracle as go


def tiny_cfg(neighbor_feats=False):
    g = cfg.gnet
    g.num_blocks = 2
    g.shortcut_dim = 8
    g.reduced_dim = 4
    g.pairfeat_dim = 6
    g.num_pwfeat_fc = 2
    g.pwfeat_dim = 7
    g.pwfeat_narrow_dim = 5
    g.predict_fc_dim = 8
    g.bias_const_init = 0.1
    g.neighbor_feats = neighbor_feats


def setup(n=14, seed=3):
    layout, total = P.param_layout(1, cfg)
    flat = P.init_flat(layout, total, cfg, seed=seed).astype(np.float64)
    # make the problem generic: random biases, non-zero everywhere
    rs = np.random.RandomState(seed)
    flat += rs.normal(0, 0.05, flat.shape)
    params = P.views(layout, flat)
    img = synthetic.make_image(n, 1, image_index=seed)
    bd = go.xyxy_to_boxdata(img['dets'])
    m = go.iou(bd, bd)
    pairs = go.neighbor_pairs(m, 0.2)
    raw = go.geometry_feats(bd, m, img['det_scores'], img['det_classes'], pairs, 1, 1.0)
    labels = (rs.uniform(0, 1, n) < 0.4).astype(np.float32)
    weights = rs.uniform(0.2, 2.0, n).astype(np.float32)
    return layout, flat, params, pairs, raw, labels, weights, n


def check(neighbor_feats, normalize, imfeats=False):
    tiny_cfg(neighbor_feats)
    cfg.train.normalize_loss = normalize
    cfg.train.loss_multiplyer = 1.7
    x0 = None
    if imfeats:      # image-feature head on 2x2x3 ROI crops, one hidden layer
        cfg.gnet.imfeats = True
        cfg.gnet.imfeat_channels, cfg.gnet.imfeat_dim = 3, 5
        cfg.imfeat_crop_height = cfg.imfeat_crop_width = 2
    layout, flat, params, pairs, raw, labels, weights, n = setup()
    if imfeats:
        x0 = np.random.RandomState(5).normal(0, 1, (n, 12))
        assert 'gnet/reduce_imfeats/fully_connected_1/weights' in layout
    grads, pred = gg.gradients(params, cfg, pairs, raw, n, labels, weights, roi_x0=x0)
    rs = np.random.RandomState(0)
    worst = 0.0
    for e in layout.values():
        for _ in range(4):
            i = e.offset + rs.randint(e.size)
            h = 1e-6
            old = flat[i]
            flat[i] = old + h
            lp = gg.data_loss(gg.forward(params, cfg, pairs, raw, n, roi_x0=x0), labels, weights, cfg)
            flat[i] = old - h
            lm = gg.data_loss(gg.forward(params, cfg, pairs, raw, n, roi_x0=x0), labels, weights, cfg)
            flat[i] = old
            num = (lp - lm) / (2 * h)
            ana = grads[e.name].reshape(-1)[i - e.offset]
            worst = max(worst, abs(num - ana) / max(1e-6, abs(num), abs(ana)))
    assert worst < 1e-5, worst


def test_gradients_match_finite_differences():
    check(neighbor_feats=False, normalize=False)


def test_gradients_with_neighbor_feats_and_normalized_loss():
    check(neighbor_feats=True, normalize=True)


def test_gradients_with_image_feature_head():
    check(neighbor_feats=False, normalize=False, imfeats=True)


def test_forward_matches_float32_oracle():
    tiny_cfg()
    layout, flat, params, pairs, raw, labels, weights, n = setup()
    p32 = dict((k, v.astype(np.float32)) for k, v in params.items())
    img = synthetic.make_image(n, 1, image_index=3)
    ref = go.gnet_forward({k: img[k] for k in ('dets', 'det_scores', 'det_classes')}, p32, cfg, 1)
    pred = gg.forward(params, cfg, pairs, raw, n)
    assert np.allclose(pred, ref['prediction'], rtol=1e-4, atol=1e-5)


def test_adam_reference_first_step_is_lr_sized():
    th, m, v = gg.adam_reference(np.zeros(3), np.array([1.0, -2.0, 0.5]), np.zeros(3), np.zeros(3),
                                 lr=0.01, step=1)
    assert np.allclose(th, [-0.01, 0.01, -0.01], atol=1e-6)
