"""ORACLE SUPPORT (test infrastructure): execute the reference's OWN model code
to produce golden vectors.

    python oracle/run_reference_graph.py            # writes tests/golden/*.npz

What runs: `/root/reference/nms_net/network.py` (class Gnet: boxdata, IoU,
tf.where neighbor build, geometry features, pair-feature MLP, blocks, predict
head, matching + loss), imported unmodified from where it lies, on
  * oracle/tf012_numpy  - an eager float32 numpy stand-in for the TF ~0.12 ops
    that file calls (TensorFlow itself cannot be installed here), and
  * oracle/_ref/libdet_matching_ref.so - the reference's own det_matching.cc
    compiled unmodified (oracle/Makefile) standing in for `det_matching.so`.
Inputs come from gossipnet_b200.synthetic / params (pure numpy generators).
The outputs are committed under tests/golden/ and pin oracle/gnet_oracle.py
(tests/test_oracle_golden.py); the GPU tests compare the CUDA path with both.
Nothing under /root/reference is copied; this script only runs in the build
container (the GPU box has no /root/reference and never needs it).
"""
import importlib
import os
import sys

import numpy as np
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get('GOSSIPNET_REFERENCE', '/root/reference')

# name, experiment yaml, num_classes, n_dets, num_blocks, image index
CASES = [
    ('person_n40_b2', 'coco_person', 1, 40, 2, 0),
    ('person_n300_b2', 'coco_person', 1, 300, 2, 0),       # BASELINE configs[0]
    ('person_n120_b16', 'coco_person', 1, 120, 16, 3),
    ('multiclass_n150_b3', 'coco_multiclass', 80, 150, 3, 1),
    ('default_rawpw_n60_b2', None, 1, 60, 2, 5),           # num_pwfeat_fc = 0 (reference default)
    # image-feature head (network.py:223-240): cfg.gnet.imfeats with a synthetic stride-16
    # feature map standing in for ResNet-101's block3/unit_22 output (C = 24 channels)
    ('imfeats_n50_b2', 'coco_person', 1, 50, 2, 7),
]
IMFEAT_CHANNELS, IMFEAT_DIM = 24, 40


def synthetic_feature_map(channels, image_index):
    """[1, 38, 63, C] float32: the stride-16 map of the 600 x 1000 synthetic canvas."""
    rs = np.random.RandomState(9000 + image_index)
    return rs.normal(0.0, 1.0, (1, 38, 63, channels)).astype(np.float32)


def _load_by_path(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def main(out_dir):
    import importlib.util  # noqa: F401
    # our own pure-numpy generators, loaded by path so that the name `nms_net`
    # stays free for the reference's package
    sys.path.insert(0, ROOT)
    from gossipnet_b200 import params as P
    from gossipnet_b200 import synthetic
    from gossipnet_b200.nms_net import config as our_config
    from oracle import det_matching_oracle
    for k in [k for k in sys.modules if k == 'nms_net' or k.startswith('nms_net.')]:
        del sys.modules[k]
    sys.path.remove(ROOT)

    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(HERE, 'tf012_numpy'))
    import tensorflow as tf
    assert 'tf012_numpy' in tf.__file__

    class _MatchingLib(object):
        @staticmethod
        def detection_matching(iou, score, ignore):
            lab, w, asg = det_matching_oracle.ref_detection_matching(
                tf._v(iou), tf._v(score), tf._v(ignore))
            return tf.T(lab), tf.T(w), tf.T(asg)

    tf.OP_LIBRARIES['det_matching.so'] = _MatchingLib()
    from oracle import roi_pool_oracle

    class _RoiLib(object):
        """roi_pooling.so stand-in: the reference's own roi_pooling_op.cc (CPU kernel),
        compiled unmodified into oracle/_ref."""
        roi_pool_grad = None

        @staticmethod
        def roi_pool(data, rois, pooled_height, pooled_width, spatial_scale):
            top, arg = roi_pool_oracle.ref_roi_pool(tf._v(data), tf._v(rois), pooled_height,
                                                    pooled_width, spatial_scale)
            return tf.T(top), tf.T(arg)

    tf.OP_LIBRARIES['roi_pooling.so'] = _RoiLib()
    if not det_matching_oracle.have_reference_build():
        det_matching_oracle.build()

    import nms_net                      # the REFERENCE package
    assert os.path.realpath(nms_net.__file__).startswith(os.path.realpath(REF)), nms_net.__file__
    from nms_net import network as ref_network
    from nms_net.config import cfg as ref_cfg, _merge_a_into_b
    from easydict import EasyDict
    import copy
    ref_defaults = copy.deepcopy(ref_cfg)

    os.makedirs(out_dir, exist_ok=True)
    for name, exp, num_classes, n_dets, num_blocks, img_idx in CASES:
        # --- configure both cfgs identically
        for k in list(ref_cfg.keys()):
            del ref_cfg[k]
        for k, v in copy.deepcopy(ref_defaults).items():
            ref_cfg[k] = v
        our_config.reset_cfg()
        if exp is not None:
            path = os.path.join(REF, 'experiments', exp, 'conf.yaml')
            with open(path) as f:
                _merge_a_into_b(EasyDict(yaml.safe_load(f)), ref_cfg)
            our_config.cfg_from_file(path)
        ref_cfg.gnet.num_blocks = num_blocks
        our_config.cfg.gnet.num_blocks = num_blocks
        with_imfeats = name.startswith('imfeats')
        if with_imfeats:
            for c in (ref_cfg, our_config.cfg):
                c.gnet.imfeats = True
                c.gnet.imfeat_dim = IMFEAT_DIM
            our_config.cfg.gnet.imfeat_channels = IMFEAT_CHANNELS

        layout, total = P.param_layout(num_classes, our_config.cfg)
        flat = P.init_flat(layout, total, our_config.cfg, seed=1000 + img_idx)
        tf.reset()
        tf.PARAMS.clear()
        tf.PARAMS.update(P.views(layout, flat))

        img = synthetic.make_image(n_dets, num_classes, seed=42, image_index=img_idx)
        cw = np.linspace(0.5, 1.5, num_classes + 1).astype(np.float32)
        batch = dict((k, tf.T(v)) for k, v in img.items())
        if with_imfeats:
            # ResNet-101 is not run: get_resnet hands back the synthetic map at stride 16
            fmap = synthetic_feature_map(IMFEAT_CHANNELS, img_idx)
            batch['image'] = tf.T(np.zeros((1, 600, 1000, 3), dtype=np.float32))
            ref_network.get_resnet = lambda image, reuse: (tf.T(fmap), 16, [], {})
        net = ref_network.Gnet(num_classes, class_weights=cw, batch=batch)

        used = sorted(v.name[:-2] for v in net.trainable_variables)
        assert used == sorted(layout.keys()), (set(used) ^ set(layout.keys()))

        out = dict(
            num_classes=np.int32(num_classes), n_dets=np.int32(n_dets),
            num_blocks=np.int32(num_blocks), image_index=np.int32(img_idx),
            param_seed=np.int32(1000 + img_idx), class_weights=cw,
            experiment=np.str_(exp or ''),
            det_det_iou=tf._v(net.det_det_iou),
            det_anno_iou=tf._v(net.det_anno_iou),
            neighbor_pair_idxs=tf._v(net.neighbor_pair_idxs),
            pw_feats=tf._v(net.pw_feats),
            block1_feats=tf._v(net.block_feats[1]),
            last_feats=tf._v(net.block_feats[-1]),
            prediction=tf._v(net.prediction),
            labels=tf._v(net.labels), weights=tf._v(net.weights),
            det_gt_matching=tf._v(net.det_gt_matching),
            loss=tf._v(net.loss), loss_normed=tf._v(net.loss_normed),
            loss_unnormed=tf._v(net.loss_unnormed))
        if with_imfeats:
            out.update(imfeats=fmap, roifeats=tf._v(net.roifeats), frcn_boxes=tf._v(net.frcn_boxes),
                       det_imfeats=tf._v(net.det_imfeats), block0_feats=tf._v(net.block_feats[0]),
                       imfeat_dim=np.int32(IMFEAT_DIM), imfeat_channels=np.int32(IMFEAT_CHANNELS))
        if n_dets > 150:   # keep the fixtures small: drop the big dense tensors
            for k in ('det_det_iou', 'pw_feats', 'block1_feats', 'last_feats'):
                out[k + '_sum'] = np.float64(np.sum(out[k], dtype=np.float64))
                del out[k]
        np.savez_compressed(os.path.join(out_dir, name + '.npz'), **out)
        print('%-24s N=%d P=%d matched=%d loss=%.6f' % (
            name, n_dets, out['neighbor_pair_idxs'].shape[0], int((out['labels'] > 0).sum()),
            float(out['loss'])))


if __name__ == '__main__':
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'tests', 'golden'))
