"""`detection_matching`: the reference's DetectionMatching op on the GPU.

Reference: nms_net/matching_module/__init__.py:10-13 (python name, loaded with
tf.load_op_library and declared NotDifferentiable) and det_matching.cc:16-33
(op definition), :72-93 (argument checks), :95-159 (algorithm).

    labels, weights, assignment = detection_matching(iou, score, ignore)

iou[N,G] float32, score[N] float32, ignore[G] bool -> labels[N] float32,
weights[N] float32, assignment[N] int32.  Inputs may be CUDA tensors (used in
place) or host arrays (copied to cuda:0); outputs are CUDA tensors.  The result
is not differentiable (outputs carry no autograd history).
"""
import numpy as np
import torch

from gossipnet_b200 import ops

__all__ = 'detection_matching'


def _dev(x, dtype):
    if not isinstance(x, torch.Tensor):
        x = torch.from_numpy(np.ascontiguousarray(x))
    return x.detach().to(device='cuda', dtype=dtype).contiguous()


def detection_matching(iou, score, ignore):
    # the reference's InvalidArgument checks (det_matching.cc:76-93)
    if np.ndim(iou) != 2:
        raise ValueError('DetectionMatching expects a 2-D vector as input 1.')
    if np.ndim(score) != 1:
        raise ValueError('DetectionMatching expects a 1-D vector as input 2.')
    if np.ndim(ignore) != 1:
        raise ValueError('DetectionMatching expects a 1-D vector as input 3.')
    n, g = int(iou.shape[0]), int(iou.shape[1])
    if n != int(score.shape[0]):
        raise ValueError('DetectionMatching expects dim 1 of input 1 and dim 1 of input 2 to be '
                         'the same (%d != %d)' % (n, int(score.shape[0])))
    if g != int(ignore.shape[0]):
        raise ValueError('DetectionMatching expects dim 2 of input 1 and dim 1 of input 3 to be '
                         'the same (%d != %d)' % (g, int(ignore.shape[0])))
    iou = _dev(iou, torch.float32)
    score = _dev(score, torch.float32)
    if isinstance(ignore, torch.Tensor):
        ignore = ignore.to(torch.uint8)
    else:
        ignore = np.asarray(ignore).astype(np.uint8)
    ignore = _dev(ignore, torch.uint8)
    dev = score.device
    img_off = torch.tensor([0, n], dtype=torch.int32, device=dev)
    gt_off = torch.tensor([0, g], dtype=torch.int32, device=dev)
    iou_off = torch.zeros(2, dtype=torch.int64, device=dev)
    return ops.detection_matching_batched(iou.view(-1), iou_off, score, ignore, img_off, gt_off, g)
