"""CPU: host-side logic that mirrors the reference's Python surface."""
import numpy as np
import pytest

from gossipnet_b200 import params as P
from gossipnet_b200 import synthetic
from gossipnet_b200.nms_net import config as C
from gossipnet_b200.nms_net.config import cfg
from tests.helpers import load_experiment


def test_cfg_defaults_match_reference_keys():
    assert cfg.gnet.neighbor_thresh == 0.2 and cfg.gnet.num_blocks == 16
    assert cfg.gnet.shortcut_dim == 128 and cfg.gnet.reduced_dim == 32
    assert cfg.gnet.pairfeat_dim == 64 and cfg.gnet.num_pwfeat_fc == 0
    assert cfg.train.normalize_loss is False and cfg.train.loss_multiplyer == 1.0


def test_cfg_merge_rules():
    with pytest.raises(KeyError):
        C.cfg_from_dict({'gnet': {'no_such_key': 1}})
    with pytest.raises(ValueError):
        C.cfg_from_dict({'gnet': {'num_blocks': '16'}})
    C.cfg_from_dict({'gnet': {'num_blocks': 4}})
    assert cfg.gnet.num_blocks == 4


def test_shipped_experiments_load():
    load_experiment('coco_person')
    assert cfg.gnet.num_blocks == 1 and cfg.gnet.num_pwfeat_fc == 3
    assert cfg.gnet.pwfeat_narrow_dim == 32 and cfg.gnet.bias_const_init == 0.1
    C.reset_cfg()
    load_experiment('coco_multiclass')
    assert cfg.gnet.num_blocks == 16 and cfg.gnet.bias_const_init == 0.01


def test_param_counts_match_survey():
    load_experiment('coco_person', num_blocks=16)
    layout, total = P.param_layout(1, cfg)
    assert P.num_params(layout) == 541345
    C.reset_cfg()
    load_experiment('coco_multiclass')
    layout, total = P.param_layout(80, cfg)
    assert P.num_params(layout) == 581793
    names = list(layout)
    assert 'gnet/pw_feats/fc1/weights' in names and 'gnet/block16/fc2/biases' in names
    assert 'gnet/predict/logits/fully_connected/weights' in names
    assert layout['gnet/block3/pw_fc1/weights'].shape == (96, 64)
    assert not layout['gnet/predict/fc1/fully_connected/weights'].regularized
    assert layout['gnet/block1/fc1/weights'].regularized
    assert not layout['gnet/block1/fc1/biases'].regularized
    assert all(e.offset % 4 == 0 for e in layout.values())


def test_synthetic_boxes_are_valid():
    img = synthetic.make_image(1000, 1)
    d = img['dets']
    assert d.dtype == np.float32 and d.shape == (1000, 4)
    assert np.all(d[:, 2] >= d[:, 0] + 4) and np.all(d[:, 3] >= d[:, 1] + 4)
    assert np.all(d >= 0) and np.all(d[:, 2] <= 1000) and np.all(d[:, 3] <= 600)
    img2 = synthetic.make_image(1000, 1)
    assert np.array_equal(d, img2['dets'])


def test_class_equal_weights():
    from gossipnet_b200.nms_net.class_weights import class_equal_weights_from_counts
    cfg.train.pos_weight = 0.1
    w = class_equal_weights_from_counts([900, 50, 50])
    assert np.allclose(w, [1000 * 0.9 / 900, 1000 * 0.05 / 50, 1000 * 0.05 / 50])


def test_nms_net_alias_package_is_the_implementation():
    import nms_net
    from nms_net import cfg as cfg2, matching_module
    from nms_net.network import Gnet
    import gossipnet_b200.nms_net.network as impl
    assert cfg2 is cfg and Gnet is impl.Gnet
    assert hasattr(matching_module, 'detection_matching')
    spec = Gnet.get_batch_spec(1, is_training=False)
    assert sorted(spec) == ['det_classes', 'det_scores', 'dets']
    assert sorted(Gnet.get_batch_spec(80)) == ['det_classes', 'det_scores', 'dets', 'gt_boxes',
                                               'gt_classes', 'gt_crowd']
    assert nms_net.cfg is cfg


def test_product_never_imports_the_oracle():
    import os
    from tests.helpers import ROOT
    for dirpath, _, files in os.walk(os.path.join(ROOT, 'gossipnet_b200')):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(dirpath, f)).read()
                assert 'import oracle' not in text and 'from oracle' not in text, f


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU restatement timed on the host cores) must print
    ONE JSON line with the same metric / unit / config keys as the GPU arm plus impl,
    cpu_baseline and a zero-copy e2e object."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, 'bench.py'), '--impl', 'reference',
                          '--steps', '1', '--warmup', '1', '--n-dets', '80', '--blocks', '2'],
                         capture_output=True, text=True, timeout=300, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['unit'] == 'detections/s' and d['higher_is_better'] is True
    assert d['metric'].startswith('detections/sec Gnet fwd')
    assert d['value'] > 0 and d['steps'] == 1 and d['n_gpus'] == 1
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1
    assert d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == {'value': d['value'], 'unit': 'detections/s', 'h2d_bytes_per_step': 0,
                        'd2h_bytes_per_step': 0}
    assert 'workload' in d['config']


def test_prefetcher_reraises_loader_errors_and_passes_feature_maps():
    """A failing loader thread must surface in get() (it used to die silently and block the
    training loop forever); roidb entries that carry their feature map pass through load_roi
    when cfg.gnet.imfeats asks for image features."""
    import numpy as np
    import pytest
    from gossipnet_b200.nms_net.dataset import Prefetcher, load_roi

    class Broken(object):
        def __init__(self):
            self.n = 0

        def next_batch(self):
            self.n += 1
            if self.n == 3:
                raise KeyError('bad roidb entry')
            return {'id': self.n}

    p = Prefetcher(Broken(), 10, q_size=2).start()
    assert p.get() == [{'id': 1}] and p.get() == [{'id': 2}]
    with pytest.raises(KeyError):
        p.get()
    p.stop()

    roi = {'dets': np.zeros((1, 4), np.float32), 'imfeats': np.zeros((1, 4, 4, 8), np.float32)}
    assert load_roi(True, roi)['imfeats'] is roi['imfeats']
    with pytest.raises(NotImplementedError):
        load_roi(True, {'dets': np.zeros((1, 4), np.float32)})
