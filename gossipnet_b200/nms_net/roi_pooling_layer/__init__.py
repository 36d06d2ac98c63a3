"""Mirror of the reference's nms_net/roi_pooling_layer package."""
