"""Shared helpers for the parity tests."""
import os

import numpy as np

from gossipnet_b200 import params as P
from gossipnet_b200.nms_net.config import cfg, cfg_from_file

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def load_experiment(name, **gnet_overrides):
    """cfg <- experiments/<name>/conf.yaml (+ overrides of cfg.gnet keys)."""
    cfg_from_file(os.path.join(ROOT, 'experiments', name, 'conf.yaml'))
    for k, v in gnet_overrides.items():
        assert k in cfg.gnet, k
        cfg.gnet[k] = v
    return cfg


def make_params(num_classes, seed=7):
    layout, total = P.param_layout(num_classes, cfg)
    flat = P.init_flat(layout, total, cfg, seed=seed)
    return layout, flat, P.views(layout, flat)


def rel_err(a, b):
    """max |a-b| / max(|b|, eps): the 1e-4 logit criterion of BASELINE.json."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), 1e-12))
