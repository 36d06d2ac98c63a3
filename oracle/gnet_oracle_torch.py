"""ORACLE (test infrastructure, NOT product code): the Gnet forward of the reference
restated with PyTorch CPU float32 ops - the stock-library CPU baseline BASELINE.md §4
names ("PyTorch-CPU float32, torch.set_num_threads(os.cpu_count())").

Only `tests/` and `bench.py`'s cpu_baseline / `--impl reference` leg import this module.
It is a second, independent statement of the same algorithm as oracle/gnet_oracle.py
(numpy); tests/test_oracle_torch.py pins it against that one and against the golden
vectors of tests/golden/ (neighbor lists identical, logits within 1e-5).  The op
sequence follows nms_net/network.py of the reference op for op: dense IoU
(:474-511), tf.where (:192-195), the gathered pair geometry (:411-454), the pair-feature
FCs (:324-342), the blocks (:344-409, tf.segment_max at :387-388) and the head (:257-273).
"""
import math

import numpy as np
import torch


def _boxdata(dets):
    """network.py:462-472."""
    x1, y1, x2, y2 = dets[:, 0], dets[:, 1], dets[:, 2], dets[:, 3]
    w = x2 - x1
    h = y2 - y1
    return x1, y1, w, h, x2, y2, w * h


def dense_iou(bd):
    """network.py:474-511 with the same box set on both sides."""
    x1, y1, _, _, x2, y2, area = bd
    iw = (torch.minimum(x2[:, None], x2[None, :]) - torch.maximum(x1[:, None], x1[None, :])).clamp_(min=0)
    ih = (torch.minimum(y2[:, None], y2[None, :]) - torch.maximum(y1[:, None], y1[None, :])).clamp_(min=0)
    inter = iw * ih
    return inter / ((area[:, None] + area[None, :]) - inter)


def pair_geometry(bd, iou, scores, classes, c, n, num_classes, mult):
    """network.py:411-454 and the multiplier at :199-200."""
    x1, y1, w, h = bd[0], bd[1], bd[2], bd[3]
    if num_classes > 1:
        sc = torch.zeros(scores.shape[0], num_classes)
        sc[torch.arange(scores.shape[0]), classes.long() - 1] = scores
    else:
        sc = scores[:, None]
    cw, ch, nw, nh = w[c], h[c], w[n], h[n]
    scale = (cw + ch) / 2
    xd = (x1[n] + nw / 2) - (x1[c] + cw / 2)
    yd = (y1[n] + nh / 2) - (y1[c] + ch / 2)
    l2 = torch.sqrt(xd * xd + yd * yd) / scale
    log2 = np.float32(math.log(2.0))
    cols = [iou[c, n], xd / scale, yd / scale, l2, torch.log(nw / cw) / log2,
            torch.log(nh / ch) / log2, (torch.log(nw / nh) - torch.log(cw / ch)) / log2]
    out = torch.cat([sc[c], sc[n]] + [v[:, None] for v in cols], dim=1)
    return out * np.float32(mult)


class TorchGnet(object):
    """Weights are taken once from the name -> ndarray views of gossipnet_b200.params
    (TF variable names and [in,out] layout of the reference)."""

    def __init__(self, params, cfg, num_classes):
        self.cfg, self.num_classes = cfg, num_classes
        self.p = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in params.items()}

    def fc(self, x, scope, relu):
        y = torch.addmm(self.p[scope + '/biases'], x, self.p[scope + '/weights'])
        return torch.relu_(y) if relu else y

    def block(self, b, infeats, c, n, not_self, seg_len, pw):
        g = self.cfg.gnet
        s = 'gnet/block%d/' % b
        feats = self.fc(infeats, s + 'reduce_dim', True)
        nfeats = self.fc(infeats, s + 'reduce_dim_neighbor', True) if g.neighbor_feats else feats
        x = torch.cat([pw, feats[c], nfeats[n] * not_self], dim=1)
        for i in range(1, g.num_block_pw_fc + 1):
            x = self.fc(x, s + 'pw_fc%d' % i, True)
        x = torch.segment_reduce(x, 'max', lengths=seg_len, axis=0)
        for i in range(1, g.num_block_fc):
            x = self.fc(x, s + 'fc%d' % i, True)
        x = self.fc(x, s + 'fc%d' % g.num_block_fc, False)
        return torch.relu_(infeats + x)

    @torch.no_grad()
    def forward(self, image, keep=False):
        """network.py:148-273 for one image without image features; returns the logits
        (and the neighbor list when `keep`)."""
        g = self.cfg.gnet
        dets = torch.from_numpy(np.ascontiguousarray(image['dets'], dtype=np.float32)).reshape(-1, 4)
        scores = torch.from_numpy(np.ascontiguousarray(image['det_scores'], dtype=np.float32))
        classes = torch.from_numpy(np.ascontiguousarray(image['det_classes'], dtype=np.int32))
        num = dets.shape[0]
        bd = _boxdata(dets)
        iou = dense_iou(bd)
        pairs = torch.nonzero(iou >= np.float32(g.neighbor_thresh))       # row-major, like tf.where
        c, n = pairs[:, 0], pairs[:, 1]
        pw = pair_geometry(bd, iou, scores, classes, c, n, self.num_classes, g.pw_feat_multiplyer)
        for i in range(1, g.num_pwfeat_fc + 1):
            pw = self.fc(pw, 'gnet/pw_feats/fc%d' % i, True)
        seg_len = torch.bincount(c, minlength=num)
        not_self = (c != n).to(torch.float32)[:, None]
        feats = torch.zeros(num, g.shortcut_dim)
        for b in range(1, g.num_blocks + 1):
            feats = self.block(b, feats, c, n, not_self, seg_len, pw)
        for i in range(1, g.num_predict_fc):
            feats = self.fc(feats, 'gnet/predict/fc%d/fully_connected' % i, False)
        logits = self.fc(feats, 'gnet/predict/logits/fully_connected', False).reshape(-1)
        if keep:
            return logits.numpy(), pairs.numpy()
        return logits.numpy()
