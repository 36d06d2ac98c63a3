"""CPU: the PyTorch-CPU restatement of the forward (oracle/gnet_oracle_torch.py, the
stock-library baseline bench.py times) against the golden vectors produced by executing
the reference's own nms_net/network.py, and against the numpy oracle."""
import numpy as np
import pytest

from gossipnet_b200 import params as P
from gossipnet_b200 import synthetic
from gossipnet_b200.nms_net.config import cfg
from oracle import gnet_oracle, gnet_oracle_torch
from tests.helpers import load_experiment, make_params, rel_err
from tests.test_oracle_golden import CASES, setup_case


@pytest.mark.parametrize('name', [c for c in CASES if not c.startswith('imfeats')])
def test_torch_restatement_matches_reference_execution(name):
    g, num_classes, layout, flat, img = setup_case(name)
    net = gnet_oracle_torch.TorchGnet(P.views(layout, flat), cfg, num_classes)
    logits, pairs = net.forward(img, keep=True)
    assert pairs.dtype == np.int64
    assert np.array_equal(pairs, g['neighbor_pair_idxs'])
    assert rel_err(logits, g['prediction']) < 1e-5


@pytest.mark.parametrize('exp,num_classes,n', [('coco_person', 1, 1), ('coco_person', 1, 500),
                                               ('coco_multiclass', 80, 200)])
def test_torch_restatement_matches_numpy_oracle(exp, num_classes, n):
    load_experiment(exp, num_blocks=3)
    _, _, pv = make_params(num_classes)
    img = synthetic.make_image(n, num_classes, image_index=5)
    ref = gnet_oracle.gnet_forward(img, pv, cfg, num_classes, keep_intermediates=False)
    logits, pairs = gnet_oracle_torch.TorchGnet(pv, cfg, num_classes).forward(img, keep=True)
    assert np.array_equal(pairs, ref['neighbor_pair_idxs'])
    assert rel_err(logits, ref['prediction']) < 1e-5
