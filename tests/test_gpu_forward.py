"""GPU: whole Gnet forward (+ matching + loss) through the nms_net surface vs
(a) the golden vectors made by executing the reference's own code and (b) the
oracle on fresh synthetic inputs.  Index outputs bit-exact; logits within the
1e-4 relative tolerance BASELINE.json states (max|d| / max|ref|)."""
import glob
import os

import numpy as np
import pytest
import torch

from gossipnet_b200 import params as P
from gossipnet_b200 import synthetic
from gossipnet_b200.nms_net.config import cfg
from gossipnet_b200.nms_net.network import Gnet
from oracle import det_matching_oracle, gnet_oracle
from tests.helpers import GOLDEN, load_experiment, rel_err
from tests.test_oracle_golden import CASES, setup_case

pytestmark = pytest.mark.gpu
LOGIT_TOL = 1e-4


def check_against(net, ref, with_loss=True):
    assert np.array_equal(net.neighbor_pair_idxs.cpu().numpy(), ref['neighbor_pair_idxs'])
    assert net.neighbor_pair_idxs.dtype == torch.int64
    pred = net.prediction.cpu().numpy()
    assert rel_err(pred, ref['prediction']) < LOGIT_TOL
    if with_loss:
        assert np.array_equal(net.det_anno_iou.cpu().numpy(), ref['det_anno_iou'])
        # matching consumes the GPU logits: identical decisions unless two logits
        # are closer than the logit tolerance (not the case for these inputs)
        assert np.array_equal(net.det_gt_matching.cpu().numpy(), ref['det_gt_matching'])
        assert np.array_equal(net.labels.cpu().numpy(), ref['labels'])
        assert np.allclose(net.weights.cpu().numpy(), ref['weights'], rtol=1e-6)
        for k in ('loss', 'loss_normed', 'loss_unnormed'):
            assert abs(float(getattr(net, k)) - float(ref[k])) <= 2e-4 * max(1.0, abs(float(ref[k])))


@pytest.mark.parametrize('name', CASES)
@pytest.mark.parametrize('fused', ['tma', 'pipe', 'ab', 'hl', False])
def test_forward_matches_reference_golden(name, fused):
    g, num_classes, layout, flat, img = setup_case(name)
    net = Gnet(num_classes, class_weights=g['class_weights'], params=flat)
    net.engine.use_fused = bool(fused)
    if fused:
        net.engine.pair_mode = fused     # both tensor-core formulations of the pair stage
    net(img)
    check_against(net, g)
    if 'pw_feats' in g:
        assert rel_err(net.pw_feats.cpu().numpy(), g['pw_feats']) < 2e-5
        assert rel_err(net.block_feats[1].cpu().numpy(), g['block1_feats']) < 2e-5
        assert rel_err(net.block_feats[-1].cpu().numpy(), g['last_feats']) < LOGIT_TOL
        assert np.array_equal(net.det_det_iou.cpu().numpy(), g['det_det_iou'])
    if 'imfeats' in g:      # image-feature head: boxes and ROI pooling bit-exact, FCs to tolerance
        assert np.array_equal(net.frcn_boxes.cpu().numpy(), g['frcn_boxes'])
        assert np.array_equal(net.roifeats.cpu().numpy(), g['roifeats'])
        assert rel_err(net.det_imfeats.cpu().numpy(), g['det_imfeats']) < 2e-5
        assert rel_err(net.block_feats[0].cpu().numpy(), g['block0_feats']) < 2e-5


@pytest.mark.parametrize('n,blocks,exp,C', [(300, 2, 'coco_person', 1),       # BASELINE configs[0]
                                            (1000, 16, 'coco_person', 1),     # configs[1]
                                            (2000, 16, 'coco_multiclass', 80)])  # configs[2] (fp32)
@pytest.mark.parametrize('mode', ['tma', 'pipe'])
def test_forward_matches_oracle_at_baseline_configs(n, blocks, exp, C, mode, oracle_built):
    load_experiment(exp, num_blocks=blocks)
    layout, total = P.param_layout(C, cfg)
    flat = P.init_flat(layout, total, cfg, seed=11)
    img = synthetic.make_image(n, C, image_index=0)
    ref = gnet_oracle.gnet_forward(img, P.views(layout, flat), cfg, C,
                                   matching_fn=det_matching_oracle.detection_matching)
    net = Gnet(C, params=flat)
    net.engine.pair_mode = mode
    net(img)
    check_against(net, ref)
    # per-logit relative error next to the max-normalised criterion (informational bound:
    # logits near zero make it arbitrarily large, so it is reported for |ref| > 0.1 max|ref|)
    pred, r = net.prediction.cpu().numpy().astype(np.float64), ref['prediction'].astype(np.float64)
    big = np.abs(r) > 0.1 * np.max(np.abs(r))
    per_elem = float(np.max(np.abs(pred - r)[big] / np.abs(r)[big]))
    print('config N=%d C=%d mode=%s: max-normalised %.2e, per-element (|ref|>10%% max) %.2e'
          % (n, C, mode, rel_err(pred, r), per_elem))
    assert per_elem < 1e-3


def test_batched_forward_equals_per_image():
    load_experiment('coco_person', num_blocks=4)
    net = Gnet(1)
    sizes = [300, 1, 1000, 57, 640]
    imgs = [synthetic.make_image(n, 1, image_index=i) for i, n in enumerate(sizes)]
    res = net.run_batch(imgs)
    off = res['img_off_host']
    batched = res['prediction'].cpu().numpy().copy()
    assign = res['det_gt_matching'].cpu().numpy().copy()
    for i, img in enumerate(imgs):
        single = net(img).cpu().numpy()
        assert np.array_equal(batched[off[i]:off[i + 1]], single)   # same kernels, same order
        assert np.array_equal(assign[off[i]:off[i + 1]], net.det_gt_matching.cpu().numpy())


def test_capacity_regrows_on_denser_batch():
    load_experiment('coco_person', num_blocks=2)
    net = Gnet(1)
    small = synthetic.make_image(100, 1)
    net(small)
    cap0 = net.engine.capacity
    big = synthetic.make_image(2000, 1)
    pred = net(big).cpu().numpy()
    assert net.engine.capacity > cap0
    layout, total = P.param_layout(1, cfg)
    ref = gnet_oracle.gnet_forward({k: v for k, v in big.items() if not k.startswith('gt_')},
                                   P.views(layout, net.engine.flat.cpu().numpy()), cfg, 1)
    assert rel_err(pred, ref['prediction']) < LOGIT_TOL


def test_inference_without_ground_truth_and_reuse():
    load_experiment('coco_person', num_blocks=2)
    net = Gnet(1)
    img = synthetic.make_image(64, 1)
    test_batch = {k: img[k] for k in Gnet.get_batch_spec(1, is_training=False)}
    pred = net(test_batch)
    assert pred.shape == (64,) and net.labels is None and net.loss is None
    val = Gnet(1, reuse=True)
    assert val.engine is net.engine
    assert torch.equal(val(test_batch), pred)
    names = [v.name for v in net.trainable_variables]
    assert 'gnet/block1/pw_fc1/weights:0' in names and len(names) == len(net.engine.layout)


def test_static_block_matches_oracle():
    load_experiment('coco_person', num_blocks=2)
    net = Gnet(1)
    img = synthetic.make_image(200, 1)
    net(img)
    pairs = net.neighbor_pair_idxs
    infeats = net.block_feats[1]
    out = Gnet._block(2, infeats, None, None, pairs[:, 0], pairs[:, 1], net.pw_feats, None)
    # Gnet._block runs the fp32 det-level layers, the full forward the fused bf16x3 ones
    assert rel_err(out.cpu().numpy(), net.block_feats[2].cpu().numpy()) < 3e-5


def test_imfeats_needs_the_feature_map():
    """cfg.gnet.imfeats: the head runs on a feature map in the batch; an `image` alone
    (ResNet-101 is not part of this package) is rejected loudly."""
    load_experiment('coco_person', num_blocks=1)
    cfg.gnet.imfeats = True
    cfg.gnet.imfeat_channels = 8
    cfg.gnet.imfeat_dim = 16
    net = Gnet(1)
    assert 'gnet/reduce_imfeats/fully_connected_1/weights' in net.state_dict()
    assert 'imfeats' in Gnet.get_batch_spec(1) and 'image' not in Gnet.get_batch_spec(1)
    img = synthetic.make_image(30, 1)
    with pytest.raises(NotImplementedError):
        net(dict(img, image=np.zeros((1, 600, 1000, 3), dtype=np.float32)))
    fmap = np.random.RandomState(0).normal(size=(1, 38, 63, 8)).astype(np.float32)
    pred = net(dict(img, imfeats=fmap))
    assert pred.shape == (30,) and net.roifeats.shape == (30, 7, 7, 8)
    assert net.det_imfeats.shape == (30, 16) and net.block_feats[0].shape == (30, 128)
    # two images in one call == one at a time
    img2 = synthetic.make_image(41, 1, image_index=3)
    fmap2 = np.random.RandomState(1).normal(size=(1, 20, 30, 8)).astype(np.float32)
    both = net.run_batch([dict(img, imfeats=fmap), dict(img2, imfeats=fmap2)])['prediction'].cpu().numpy().copy()
    assert np.array_equal(both[:30], pred.cpu().numpy())
    assert np.array_equal(both[30:], net(dict(img2, imfeats=fmap2)).cpu().numpy())


def test_collapsed_predict_head_matches_staged_layers():
    """The linear predict head folded into one affine map (gn_predict_collapse +
    gn_rowdot_fwd) against the three staged FC launches (network.py:257-273)."""
    load_experiment('coco_person', num_blocks=3)
    net = Gnet(1)
    img = synthetic.make_image(500, 1, image_index=2)
    net.engine.collapse_predict = True
    folded = net(img).cpu().numpy().copy()
    net.engine.collapse_predict = False
    staged = net(img).cpu().numpy()
    assert rel_err(folded, staged) < 5e-6


@pytest.mark.parametrize('n,blocks,exp,C', [(1000, 16, 'coco_person', 1),
                                            (2000, 16, 'coco_multiclass', 80)])   # configs[2] as named
def test_bf16_arithmetic_mode(n, blocks, exp, C, oracle_built):
    """cfg.gnet.compute_dtype = 'bf16' (BASELINE configs[2]: "bf16"): plain bf16 operands with
    fp32 accumulation in the three fused tensor-core kernels.  Index outputs do not depend on
    the mode (bit-exact); logits agree with the fp32 oracle to the stated bf16 tolerance."""
    BF16_TOL = 3e-2      # max|d| / max|ref| after 16 blocks of bf16-operand GEMMs
    load_experiment(exp, num_blocks=blocks)
    cfg.gnet.compute_dtype = 'bf16'
    layout, total = P.param_layout(C, cfg)
    flat = P.init_flat(layout, total, cfg, seed=11)
    img = synthetic.make_image(n, C, image_index=0)
    test_img = {k: img[k] for k in ('dets', 'det_scores', 'det_classes')}
    ref = gnet_oracle.gnet_forward(test_img, P.views(layout, flat), cfg, C)
    net = Gnet(C, params=flat)
    assert net.engine.bf16
    pred = net(test_img).cpu().numpy()
    assert np.array_equal(net.neighbor_pair_idxs.cpu().numpy(), ref['neighbor_pair_idxs'])
    err = rel_err(pred, ref['prediction'])
    assert 1e-6 < err < BF16_TOL, err        # really the reduced-precision path, and close
    # the same network in the default mode is 100x closer
    cfg.gnet.compute_dtype = 'fp32'
    net32 = Gnet(C, params=flat)
    assert rel_err(net32(test_img).cpu().numpy(), ref['prediction']) < LOGIT_TOL


def test_weight_derived_buffers_follow_weight_changes():
    """Operand images and the folded predict head are cached across forwards; an in-place
    change of the weights (TF-style variable assignment, a checkpoint load, an optimizer step)
    must rebuild them - in eager forwards and before the replay of a captured graph."""
    from gossipnet_b200.session import InferenceSession
    from gossipnet_b200.trainer import Trainer
    load_experiment('coco_person', num_blocks=3)
    img = synthetic.make_image(300, 1, image_index=5)
    test_batch = {k: img[k] for k in ('dets', 'det_scores', 'det_classes')}
    net = Gnet(1)
    p0 = net(test_batch).clone()
    assert torch.equal(net(test_batch), p0)                      # cached path: same result
    var = [v for v in net.trainable_variables if v.op_name == 'gnet/block2/pw_fc2/weights'][0]
    var.value.mul_(1.5)                                          # in-place assignment
    head = [v for v in net.trainable_variables if v.op_name == 'gnet/predict/logits/fully_connected/weights'][0]
    head.value.mul_(0.5)
    p1 = net(test_batch).clone()
    fresh = Gnet(1, params=net.engine.flat.cpu().numpy())        # same weights, nothing cached
    assert not torch.equal(p1, p0) and torch.equal(fresh(test_batch), p1)
    # captured graph: two runs capture it, then the weights move
    sess = InferenceSession(net)
    off = np.array([0, 300], np.int32)
    args = (img['dets'], img['det_scores'], img['det_classes'], off)
    for _ in range(3):
        g1 = sess.run(*args).copy()
    assert sess._graph and np.array_equal(g1, p1.cpu().numpy())
    tr = Trainer(net)
    tr.step([img], 1e-2)                                         # optimizer writes through the C ABI
    g2 = sess.run(*args).copy()
    assert not np.array_equal(g2, g1)
    assert np.array_equal(g2, Gnet(1, params=net.engine.flat.cpu().numpy())(test_batch).cpu().numpy())


def test_launch_chain_switches_do_not_change_the_logits():
    """Programmatic dependent launch on / off (gn_set_pdl) and the two detection-level kernels
    (copy-engine kernel of gn_det_tma.cu / eight-warp kernel of gn_det_tc.cu) are pure
    scheduling / data-movement choices: bit-identical logits, eager and as a replayed graph."""
    from gossipnet_b200 import ops
    from gossipnet_b200.session import InferenceSession
    load_experiment('coco_person', num_blocks=4)
    imgs = [synthetic.make_image(n, 1, image_index=i) for i, n in enumerate((700, 1, 333))]
    dets = np.concatenate([im['dets'] for im in imgs]).astype(np.float32)
    scores = np.concatenate([im['det_scores'] for im in imgs]).astype(np.float32)
    classes = np.concatenate([im['det_classes'] for im in imgs]).astype(np.int32)
    off = np.array([0, 700, 701, 1034], np.int32)
    results = []
    try:
        for pdl in (True, False):
            for det_tma in (True, False):
                ops.set_pdl(pdl)
                net = Gnet(1)
                net.engine.det_tma = det_tma
                sess = InferenceSession(net)
                outs = [sess.run(dets, scores, classes, off).copy() for _ in range(3)]   # third run replays the graph
                assert sess._graph
                assert np.array_equal(outs[0], outs[2])
                results.append(outs[2])
    finally:
        ops.set_pdl(True)
    for r in results[1:]:
        assert np.array_equal(r, results[0])
