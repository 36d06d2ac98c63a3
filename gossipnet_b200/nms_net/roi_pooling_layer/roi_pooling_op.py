"""`roi_pool` / `roi_pool_grad` with the reference's op signatures
(nms_net/roi_pooling_layer/roi_pooling_op.py:4-7; op definitions
roi_pooling_op.cc:35-54).  Implemented by gn_roi_pool_fwd / gn_roi_pool_bwd."""
from gossipnet_b200 import ops


def roi_pool(bottom_data, bottom_rois, pooled_height, pooled_width, spatial_scale):
    """bottom_data[B,H,W,C] f32, bottom_rois[R,5] f32 (batch_idx,x1,y1,x2,y2) ->
    (top_data[R,PH,PW,C] f32, argmax[R,PH,PW,C] i32)."""
    return ops.roi_pool_fwd(bottom_data, bottom_rois, pooled_height, pooled_width,
                            spatial_scale)


def roi_pool_grad(bottom_data, bottom_rois, argmax, grad, pooled_height, pooled_width,
                  spatial_scale):
    """-> d loss / d bottom_data [B,H,W,C] f32."""
    return ops.roi_pool_bwd(bottom_data, bottom_rois, argmax, grad, pooled_height,
                            pooled_width, spatial_scale)
