"""CPU: the oracle restatement (oracle/gnet_oracle.py) against golden vectors
produced by executing the reference's own nms_net/network.py + det_matching.cc
(oracle/run_reference_graph.py).  This is what pins the oracle."""
import glob
import os

import numpy as np
import pytest

from gossipnet_b200 import params as P
from gossipnet_b200 import synthetic
from gossipnet_b200.nms_net.config import cfg
from oracle import det_matching_oracle, gnet_oracle
from tests.helpers import GOLDEN, load_experiment, rel_err

CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, '*.npz')))


def setup_case(name):
    g = np.load(os.path.join(GOLDEN, name + '.npz'))
    exp = str(g['experiment'])
    if exp:
        load_experiment(exp)
    cfg.gnet.num_blocks = int(g['num_blocks'])
    if 'imfeats' in g:          # image-feature head: synthetic stride-16 map in the fixture
        cfg.gnet.imfeats = True
        cfg.gnet.imfeat_dim = int(g['imfeat_dim'])
        cfg.gnet.imfeat_channels = int(g['imfeat_channels'])
    num_classes = int(g['num_classes'])
    layout, total = P.param_layout(num_classes, cfg)
    flat = P.init_flat(layout, total, cfg, seed=int(g['param_seed']))
    img = synthetic.make_image(int(g['n_dets']), num_classes, seed=42,
                               image_index=int(g['image_index']))
    if 'imfeats' in g:
        img['imfeats'] = g['imfeats']
    return g, num_classes, layout, flat, img


def test_golden_files_present():
    assert len(CASES) >= 5


@pytest.mark.parametrize('name', CASES)
def test_oracle_matches_reference_execution(name, oracle_built):
    g, num_classes, layout, flat, img = setup_case(name)
    out = gnet_oracle.gnet_forward(img, P.views(layout, flat), cfg, num_classes,
                                   matching_fn=det_matching_oracle.detection_matching,
                                   class_weights=g['class_weights'])
    # integer / index outputs: bit-exact
    assert np.array_equal(out['neighbor_pair_idxs'], g['neighbor_pair_idxs'])
    assert out['neighbor_pair_idxs'].dtype == np.int64
    assert np.array_equal(out['det_gt_matching'], g['det_gt_matching'])
    assert np.array_equal(out['labels'], g['labels'])
    assert np.array_equal(out['weights'], g['weights'])
    # float tensors: same numpy float32 arithmetic -> identical up to BLAS blocking
    assert np.array_equal(out['det_anno_iou'], g['det_anno_iou'])
    if 'det_det_iou' in g:
        assert np.array_equal(out['det_det_iou'], g['det_det_iou'])
        assert rel_err(out['pw_feats'], g['pw_feats']) < 1e-6
        assert rel_err(out['block_feats'][1], g['block1_feats']) < 1e-6
        assert rel_err(out['block_feats'][-1], g['last_feats']) < 1e-6
    else:
        assert abs(np.sum(out['det_det_iou'], dtype=np.float64) - g['det_det_iou_sum']) < 1e-6
    if 'imfeats' in g:
        assert np.array_equal(out['frcn_boxes'], g['frcn_boxes'])
        assert np.array_equal(out['roifeats'], g['roifeats'])       # C restatement == reference op
        assert rel_err(out['det_imfeats'], g['det_imfeats']) < 1e-6
        assert rel_err(out['block_feats'][0], g['block0_feats']) < 1e-6
    assert rel_err(out['prediction'], g['prediction']) < 1e-6
    for k in ('loss', 'loss_normed', 'loss_unnormed'):
        assert abs(float(out[k]) - float(g[k])) <= 1e-6 * max(1.0, abs(float(g[k]))), k
