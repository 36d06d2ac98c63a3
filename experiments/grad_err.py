"""Per-parameter gradient error of the CUDA training step against the float64 oracle, with
the tensor-core FC kernels (gn_fc_tc.cu) and with the fp32 CUDA-core ones."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gossipnet_b200 import params as P, synthetic
from gossipnet_b200.nms_net.config import cfg
from gossipnet_b200.nms_net.network import Gnet
from gossipnet_b200.trainer import Trainer
from tests.helpers import load_experiment
from tests.test_gpu_training import oracle_grad_for

load_experiment('coco_person', num_blocks=2)
img = synthetic.make_image(300, 1, image_index=0)
layout, total = P.param_layout(1, cfg)
flat = P.init_flat(layout, total, cfg, seed=21)
cw = np.linspace(0.5, 1.5, 2).astype(np.float32)
got = {}


class Mixed(Trainer):
    """forward and backward FCs switchable independently"""
    tc_fwd = tc_bwd = True

    def _fc(self, *a, **k):
        self.use_tc = self.tc_fwd
        return Trainer._fc(self, *a, **k)

    def _fc_bwd(self, *a, **k):
        self.use_tc = self.tc_bwd
        return Trainer._fc_bwd(self, *a, **k)


for tc in (True, False, 'fwd', 'bwd'):
    net = Gnet(1, class_weights=cw, params=flat)
    tr = Mixed(net)
    tr.tc_fwd = tc in (True, 'fwd')
    tr.tc_bwd = tc in (True, 'bwd')
    tr.use_tc = True
    res = tr.forward_backward([img])
    got[tc] = tr.grad.cpu().numpy().astype(np.float64)
want = oracle_grad_for(img, flat, layout, 1, res['labels'].cpu().numpy(), res['weights'].cpu().numpy())
print('%-40s %10s %10s %10s %10s %10s' % ('parameter', 'tc vs f64', 'ffma vs f64', 'tc vs ffma', 'tc-fwd only', 'tc-bwd only'))
for e in layout.values():
    s = slice(e.offset, e.offset + e.size)
    sc = max(np.max(np.abs(want[s])), 1e-6)
    print('%-40s %10.2e %10.2e %10.2e %10.2e %10.2e' % (e.name, np.max(np.abs(got[True][s] - want[s])) / sc,
                                          np.max(np.abs(got[False][s] - want[s])) / sc,
                                          np.max(np.abs(got[True][s] - got[False][s])) / sc,
                                          np.max(np.abs(got['fwd'][s] - want[s])) / sc,
                                          np.max(np.abs(got['bwd'][s] - want[s])) / sc))
