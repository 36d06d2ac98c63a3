"""GPU: the training step (A10/A11) - CUDA gradients vs the float64 gradient oracle,
the optimizer updates vs their closed forms, and a few descending steps."""
import numpy as np
import pytest
import torch

from gossipnet_b200 import ops
from gossipnet_b200 import params as P
from gossipnet_b200 import synthetic
from gossipnet_b200.nms_net.config import cfg
from gossipnet_b200.nms_net.network import Gnet
from gossipnet_b200.trainer import LearningRate, Trainer
from oracle import gnet_grad_oracle as gg
from oracle import gnet_oracle as go
from tests.helpers import load_experiment

pytestmark = pytest.mark.gpu


def oracle_grad_for(img, flat, layout, num_classes, labels, weights):
    bd = go.xyxy_to_boxdata(img['dets'])
    x0 = None
    if cfg.gnet.imfeats:     # ROI-pooled crops are constants of the step
        _, roifeats, _, _ = go.image_features(bd, img['imfeats'], P.views(layout, flat), cfg)
        x0 = roifeats.reshape(roifeats.shape[0], -1)
    m = go.iou(bd, bd)
    pairs = go.neighbor_pairs(m, cfg.gnet.neighbor_thresh)
    raw = go.geometry_feats(bd, m, img['det_scores'], img['det_classes'], pairs, num_classes,
                            cfg.gnet.pw_feat_multiplyer)
    grads, _ = gg.gradients(P.views(layout, flat.astype(np.float64)), cfg, pairs, raw,
                            img['dets'].shape[0], labels, weights, roi_x0=x0)
    out = np.zeros(flat.shape[0])
    for e in layout.values():
        out[e.offset:e.offset + e.size] = grads[e.name].reshape(-1)
    return out


# Trainer arithmetic -> tolerance against the float64 gradient oracle (max |d| / max |ref| per
# parameter).  'ffma': fp32 CUDA-core FCs throughout.  'tc_bwd': fp32 forward, tensor-core
# (bf16x3) backward GEMMs on the same activations - the kernel-level check of gn_fc_tc.cu.
# 'tc' (the shipped default): tensor-core forward too; its ~1e-5 activation differences flip a
# few relu / segment-max selections, and the gradient of a piecewise-linear network jumps
# there, so the bound is the looser one (experiments/grad_err.py separates the two effects).
MODES = {'ffma': 2e-3, 'tc_bwd': 1e-4, 'tc': 2e-2}


def grad_check(num_classes, imgs, mode='tc'):
    tol = MODES[mode]
    layout, total = P.param_layout(num_classes, cfg)
    flat = P.init_flat(layout, total, cfg, seed=21)
    cw = np.linspace(0.5, 1.5, num_classes + 1).astype(np.float32)
    net = Gnet(num_classes, class_weights=cw, params=flat)
    tr = Trainer(net)
    tr.use_tc_fwd = mode == 'tc'
    tr.use_tc_bwd = mode in ('tc', 'tc_bwd')
    res = tr.forward_backward(imgs)
    got = tr.grad.cpu().numpy().astype(np.float64)
    assert float(tr.gradbuf[-1]) == len(imgs)
    off = res['img_off_host'] if 'img_off_host' in res else np.cumsum([0] + [i['dets'].shape[0] for i in imgs])
    labels = res['labels'].cpu().numpy()
    weights = res['weights'].cpu().numpy()        # already class-weighted (constants of the step)
    want = np.zeros(total)
    for i, img in enumerate(imgs):
        s = slice(int(off[i]), int(off[i + 1]))
        want += oracle_grad_for(img, flat, layout, num_classes, labels[s], weights[s])
    for e in layout.values():
        a, b = got[e.offset:e.offset + e.size], want[e.offset:e.offset + e.size]
        scale = max(np.max(np.abs(b)), 1e-6)
        assert np.max(np.abs(a - b)) / scale < tol, (e.name, np.max(np.abs(a - b)) / scale)
    return tr, res


@pytest.mark.parametrize('mode', ['tc_bwd', 'tc'])
def test_gradients_with_image_feature_head(mode, oracle_built):
    load_experiment('coco_person', num_blocks=2)
    cfg.gnet.imfeats = True
    cfg.gnet.imfeat_channels, cfg.gnet.imfeat_dim = 16, 48
    imgs = []
    for i, n in enumerate((120, 75)):
        img = synthetic.make_image(n, 1, image_index=i)
        img['imfeats'] = np.random.RandomState(40 + i).normal(size=(1, 38, 63, 16)).astype(np.float32)
        imgs.append(img)
    grad_check(1, imgs, mode)


@pytest.mark.parametrize('mode', sorted(MODES))
def test_gradients_coco_person_two_blocks(mode):
    load_experiment('coco_person', num_blocks=2)
    grad_check(1, [synthetic.make_image(300, 1, image_index=0)], mode)


@pytest.mark.parametrize('mode', sorted(MODES))
def test_gradients_multi_image_batch_sum(mode):
    load_experiment('coco_person', num_blocks=3)
    grad_check(1, [synthetic.make_image(n, 1, image_index=i) for i, n in enumerate([120, 57, 200])],
               mode)


@pytest.mark.parametrize('mode', ['tc_bwd', 'tc'])
def test_gradients_multiclass_and_normalized_loss(mode):
    load_experiment('coco_multiclass', num_blocks=2)
    cfg.train.normalize_loss = True
    cfg.train.loss_multiplyer = 3.0
    grad_check(80, [synthetic.make_image(150, 80, image_index=2)], mode)


@pytest.mark.parametrize('mode', ['tc_bwd', 'tc'])
def test_gradients_neighbor_feats_and_raw_pair_features(mode):
    cfg.gnet.num_blocks = 2
    cfg.gnet.neighbor_feats = True          # reduce_dim_neighbor branch, num_pwfeat_fc = 0
    cfg.gnet.bias_const_init = 0.05
    grad_check(1, [synthetic.make_image(90, 1, image_index=4)], mode)


def test_adam_and_momentum_updates_closed_form():
    rs = np.random.RandomState(0)
    n = 1000
    theta = rs.normal(0, 1, n).astype(np.float32)
    grad = rs.normal(0, 1, n).astype(np.float32)
    decay = (rs.uniform(0, 1, n) < 0.5).astype(np.float32) * 0.0005
    d = lambda a: torch.from_numpy(a.copy()).cuda()
    th, m, v = d(theta), torch.zeros(n, device='cuda'), torch.zeros(n, device='cuda')
    ref_t, ref_m, ref_v = theta.astype(np.float64), np.zeros(n), np.zeros(n)
    for step in (1, 2, 3):
        ops.adam_step(th, d(grad), m, v, d(decay), 1e-3, 0.9, 0.999, 1e-8, step, 0.25)
        g = 0.25 * grad + decay * ref_t
        ref_t, ref_m, ref_v = gg.adam_reference(ref_t, g, ref_m, ref_v, 1e-3, step)
        assert np.allclose(th.cpu().numpy(), ref_t, rtol=1e-5, atol=1e-6)
    th, acc = d(theta), torch.zeros(n, device='cuda')
    ops.momentum_step(th, d(grad), acc, None, 0.1, 0.9, 1.0)
    ops.momentum_step(th, d(grad), acc, None, 0.1, 0.9, 1.0)
    assert np.allclose(th.cpu().numpy(), theta - 0.1 * grad - 0.1 * 1.9 * grad, rtol=1e-5, atol=1e-6)


def test_training_steps_reduce_the_loss():
    load_experiment('coco_person', num_blocks=2)
    imgs = [synthetic.make_image(200, 1, image_index=i) for i in range(4)]
    net = Gnet(1)
    tr = Trainer(net, optimizer='adam')
    losses = []
    for it in range(12):
        res = tr.step(imgs, 1e-3)
        losses.append(float(res['loss_out'][:, 2].sum()))
    assert res['images_in_step'] == 4 and tr.global_step == 12
    assert losses[-1] < 0.8 * losses[0], losses
    sd = tr.state_dict()
    tr2 = Trainer(Gnet(1))
    tr2.load_state_dict(sd)
    assert tr2.global_step == 12 and torch.equal(tr2.eng.flat.cpu(), sd['params'])


def test_learning_rate_schedule_matches_reference_semantics():
    cfg.train.lr_multi_step = [[3, 0.1], [5, 0.01]]
    lr = LearningRate()
    got = [lr.get_lr(i) for i in range(1, 8)]
    assert got == [0.1, 0.1, 0.1, 0.01, 0.01, 0.01, 0.01]


def test_clip_gradients_per_variable_norm():
    """gn_clip_gradients = slim's clip_gradient_norm (reference train.py:73-76): every entry's
    gradient of the total loss, scale * grad + decay * theta, times clip / max(||.||, clip)."""
    rs = np.random.RandomState(3)
    sizes = [1, 7, 256, 4096, 65536, 33]
    offs = np.concatenate([[0], np.cumsum([(s + 3) // 4 * 4 for s in sizes])]).astype(np.int64)
    total = int(offs[-1])
    grad = rs.normal(0, 1, total).astype(np.float32)
    theta = rs.normal(0, 1, total).astype(np.float32)
    decay = np.where(rs.rand(total) < 0.5, 1e-2, 0.0).astype(np.float32)
    grad[offs[2]:offs[2] + sizes[2]] *= 1e-4          # an entry whose norm stays under the clip
    table = np.stack([offs[:-1], sizes], axis=1).astype(np.int32)
    scale, clip = 0.125, 2.0
    ref = grad.astype(np.float64).copy()
    for o, n in table:
        g = scale * grad[o:o + n].astype(np.float64) + decay[o:o + n].astype(np.float64) * theta[o:o + n]
        ref[o:o + n] = g * clip / max(np.sqrt(np.sum(g * g)), clip)
    d = lambda a: torch.from_numpy(a).cuda()
    g_dev = d(grad.copy())
    ops.clip_gradients(g_dev, d(theta), d(decay), d(table), scale, clip)
    got = g_dev.cpu().numpy()
    for o, n in table:
        assert np.allclose(got[o:o + n], ref[o:o + n], rtol=2e-5, atol=1e-7)
        assert np.sqrt(np.sum(got[o:o + n].astype(np.float64) ** 2)) <= clip * (1 + 1e-5)
    pad = np.ones(total, bool)
    for o, n in table:
        pad[o:o + n] = False
    assert np.array_equal(got[pad], grad[pad])         # alignment padding between entries untouched


def test_trainer_gradient_clipping():
    """cfg.train.gradient_clipping: a norm nobody reaches leaves the step unchanged; a small one
    bounds every variable's Adam input (first step: |update| = lr for every touched element, so
    compare the second moments instead)."""
    load_experiment('coco_person', num_blocks=2)
    imgs = [synthetic.make_image(120, 1, image_index=i) for i in range(2)]
    v2, init = [], None
    for clip in (-1.0, 1e9, 1e-3):
        cfg.train.gradient_clipping = clip
        try:
            tr = Trainer(Gnet(1))
            if init is None:
                init = tr.eng.flat.clone()
            tr.eng.flat.copy_(init)               # the same starting point for the three steps
            tr.eng.weights_version += 1
            tr.step(imgs, 1e-3)
            v2.append(tr.state2.clone())
        finally:
            cfg.train.gradient_clipping = -1.0
    # the gradient itself is summed with atomics (order varies run to run): compare |g| recovered
    # from Adam's second moment, (1 - beta2) g^2, in the l2 norm instead of the +-lr first-step update
    g0, g1 = torch.sqrt(v2[0] / (1 - 0.999)), torch.sqrt(v2[1] / (1 - 0.999))
    assert float((g0 - g1).norm() / g0.norm()) < 1e-5
    # Adam's v after one step = (1 - beta2) g^2: per entry sum(g^2) <= clip^2 once clipped
    tr_layout = tr.eng.layout
    for e in tr_layout.values():
        gsq = float(v2[2][e.offset:e.offset + e.size].sum()) / (1 - 0.999)
        assert gsq <= (1e-3) ** 2 * (1 + 1e-3), e.name
    assert float((v2[0] / (1 - 0.999)).sum()) > 1e-4     # the unclipped step was far above that
