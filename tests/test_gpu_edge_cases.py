"""GPU: edge cases and full-size properties of the whole path - empty and one-detection
images inside a batch, an all-isolated image (every detection only has its self pair), the
N = 10 000 stress configuration (BASELINE configs[4]) through size-independent properties."""
import numpy as np
import pytest
import torch

from gossipnet_b200 import ops
from gossipnet_b200 import params as P
from gossipnet_b200 import synthetic
from gossipnet_b200.nms_net.config import cfg
from gossipnet_b200.nms_net.network import Gnet
from oracle import det_matching_oracle, gnet_oracle
from tests.helpers import load_experiment, rel_err

pytestmark = pytest.mark.gpu
KEYS = ('dets', 'det_scores', 'det_classes')


def empty_image():
    return {'dets': np.zeros((0, 4), np.float32), 'det_scores': np.zeros((0,), np.float32),
            'det_classes': np.zeros((0,), np.int32)}


def test_batch_with_empty_and_single_detection_images(oracle_built):
    load_experiment('coco_person', num_blocks=3)
    net = Gnet(1)
    layout, _ = P.param_layout(1, cfg)
    pv = P.views(layout, net.engine.flat.cpu().numpy())
    imgs = [empty_image(),
            {k: synthetic.make_image(1, 1, image_index=4)[k] for k in KEYS},
            {k: synthetic.make_image(130, 1, image_index=5)[k] for k in KEYS},
            empty_image(),
            {k: synthetic.make_image(2, 1, image_index=6)[k] for k in KEYS}]
    res = net.run_batch(imgs)
    off = res['img_off_host']
    assert off.tolist() == [0, 0, 1, 131, 131, 133]
    pred = res['prediction'].cpu().numpy()
    for i, img in enumerate(imgs):
        if img['dets'].shape[0] == 0:
            continue
        ref = gnet_oracle.gnet_forward(img, pv, cfg, 1)
        assert rel_err(pred[off[i]:off[i + 1]], ref['prediction']) < 1e-4


def test_only_empty_images():
    load_experiment('coco_person', num_blocks=2)
    net = Gnet(1)
    res = net.run_batch([empty_image(), empty_image()])
    assert res['prediction'].numel() == 0 and res['P'] == 0


def test_isolated_detections_only_self_pairs(oracle_built):
    """Disjoint boxes: P == N, every neighbour row is zeroed (network.py:372-374), and the
    whole network reduces to the same per-detection function of (score, self features)."""
    load_experiment('coco_person', num_blocks=4)
    n = 200
    xs = (np.arange(n) % 20) * 45.0
    ys = (np.arange(n) // 20) * 50.0
    dets = np.stack([xs, ys, xs + 30.0, ys + 40.0], axis=1).astype(np.float32)
    img = {'dets': dets, 'det_scores': np.linspace(0.01, 0.99, n).astype(np.float32),
           'det_classes': np.ones(n, np.int32)}
    net = Gnet(1)
    pred = net(img).cpu().numpy()
    pairs = net.neighbor_pair_idxs.cpu().numpy()
    assert pairs.shape == (n, 2) and np.array_equal(pairs[:, 0], pairs[:, 1])
    layout, _ = P.param_layout(1, cfg)
    ref = gnet_oracle.gnet_forward(img, P.views(layout, net.engine.flat.cpu().numpy()), cfg, 1)
    assert rel_err(pred, ref['prediction']) < 1e-4
    # identical boxes up to translation and equal scores -> equal logits
    img['det_scores'][:] = 0.5
    same = net(img).cpu().numpy()
    assert np.ptp(same) <= 1e-5 * max(1.0, abs(float(same[0])))


def test_stress_n10000_properties():
    """configs[4]: N = 10 000 detections in ONE image (1.7 M pairs).  The CSR neighbor build
    (IoU recomputed in registers) must equal `where(dense_iou >= thresh)` of the dense kernel
    in row-major order; the graph is symmetric; logits are finite and do not depend on what
    else shares the batch."""
    load_experiment('coco_person', num_blocks=2)
    net = Gnet(1)
    img = {k: synthetic.make_image(10000, 1, image_index=0)[k] for k in KEYS}
    pred = net(img).cpu().numpy().copy()
    pairs = net.neighbor_pair_idxs
    assert 1600000 < pairs.shape[0] < 1800000     # ~1.7 M pairs for this recipe
    dense = ops.iou_dense(net.dets, net.dets)
    want = torch.nonzero(dense >= torch.tensor(cfg.gnet.neighbor_thresh, dtype=torch.float32,
                                               device=dense.device))
    assert torch.equal(pairs, want)           # bit-exact index lists, tf.where order
    assert torch.equal(dense, dense.t())      # bitwise symmetric
    assert bool(torch.all(torch.diagonal(dense) == 1.0))
    del dense, want
    assert np.all(np.isfinite(pred))
    other = {k: synthetic.make_image(700, 1, image_index=9)[k] for k in KEYS}
    both = net.run_batch([other, img])
    off = both['img_off_host']
    assert np.array_equal(both['prediction'].cpu().numpy()[off[1]:off[2]], pred)


@pytest.mark.parametrize('mode', ['tma', 'pipe'])
def test_stress_n10000_against_oracle(mode):
    """configs[4] against the CPU oracle itself (numpy handles 10 000^2 fp32): pair index
    lists and the pairs' IoU bits identical, logits within 1e-4 (network.py:474-511, 192-195)."""
    load_experiment('coco_person', num_blocks=2)
    net = Gnet(1)
    net.engine.pair_mode = mode
    img = {k: synthetic.make_image(10000, 1, image_index=0)[k] for k in KEYS}
    pred = net(img).cpu().numpy().copy()
    layout, _ = P.param_layout(1, cfg)
    ref = gnet_oracle.gnet_forward(img, P.views(layout, net.engine.flat.cpu().numpy()), cfg, 1)
    assert np.array_equal(net.neighbor_pair_idxs.cpu().numpy(), ref['neighbor_pair_idxs'])
    Pn = ref['neighbor_pair_idxs'].shape[0]
    d = img['dets']
    bd = gnet_oracle.xyxy_to_boxdata(d)
    c, n = ref['neighbor_pair_idxs'][:, 0], ref['neighbor_pair_idxs'][:, 1]
    want_iou = gnet_oracle.iou(bd, bd)[c, n]
    got_iou = net._res['pair_iou'][:Pn].cpu().numpy()
    assert np.array_equal(got_iou.view(np.uint32), want_iou.view(np.uint32))
    assert rel_err(pred, ref['prediction']) < 1e-4


def test_multiclass_matching_respects_classes(oracle_built):
    """Multi-class det/GT overlaps are zeroed across classes (network.py:177-187), so a
    detection can only match a ground truth of its own class."""
    load_experiment('coco_multiclass', num_blocks=1)
    net = Gnet(80)
    img = synthetic.make_image(400, 80, image_index=2)
    net(img)
    assign = net.det_gt_matching.cpu().numpy()
    matched = assign >= 0
    assert matched.any()
    assert np.array_equal(img['gt_classes'][assign[matched]], img['det_classes'][matched])
    layout, _ = P.param_layout(80, cfg)
    ref = gnet_oracle.gnet_forward(img, P.views(layout, net.engine.flat.cpu().numpy()), cfg, 80,
                                   matching_fn=det_matching_oracle.detection_matching)
    assert np.array_equal(assign, ref['det_gt_matching'])
