"""Gradient wiring of RoiPool (reference roi_pooling_op_grad.py:23-43): the
gradient w.r.t. the feature map is `roi_pool_grad(data, rois, argmax, grad,
...)`, none for the rois.  Exposed as a torch.autograd.Function."""
import torch

from gossipnet_b200.nms_net.roi_pooling_layer import roi_pooling_op


class RoiPoolFunction(torch.autograd.Function):

    @staticmethod
    def forward(ctx, data, rois, pooled_height, pooled_width, spatial_scale):
        top, argmax = roi_pooling_op.roi_pool(data, rois, pooled_height, pooled_width,
                                              spatial_scale)
        ctx.save_for_backward(data, rois, argmax)
        ctx.attrs = (pooled_height, pooled_width, spatial_scale)
        ctx.mark_non_differentiable(argmax)
        return top, argmax

    @staticmethod
    def backward(ctx, grad, _):
        data, rois, argmax = ctx.saved_tensors
        ph, pw, scale = ctx.attrs
        data_grad = roi_pooling_op.roi_pool_grad(data, rois, argmax, grad.contiguous(), ph, pw,
                                                 scale)
        return data_grad, None, None, None, None


def roi_pool_with_grad(data, rois, pooled_height, pooled_width, spatial_scale):
    return RoiPoolFunction.apply(data, rois, pooled_height, pooled_width, spatial_scale)
