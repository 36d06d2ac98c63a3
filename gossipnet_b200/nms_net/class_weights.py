"""Per-class loss weights (reference nms_net/class_weights.py:12-23).

The expected weight mass is `1 - pos_weight` for background and `pos_weight`
spread evenly over the foreground classes; index 0 is background.  The
reference counts classes through imdb.tools.get_class_counts (data plumbing,
out of scope), so the counts are an argument here.
"""
import numpy as np

from gossipnet_b200.nms_net.config import cfg


def class_equal_weights_from_counts(class_counts, num_classes=None):
    class_counts = np.asarray(class_counts, dtype=np.float64)
    if num_classes is None:
        num_classes = class_counts.shape[0] - 1
    posweight = cfg.train.pos_weight
    expected_class_weight = np.array(
        [1 - posweight] + [posweight / num_classes] * num_classes, dtype=np.float32)
    num_samples = np.sum(class_counts)
    return (num_samples * expected_class_weight / class_counts).astype(np.float32)


def class_equal_weights(imdb):
    """Same call as the reference when the imdb dict carries `class_counts`."""
    return class_equal_weights_from_counts(imdb['class_counts'], imdb['num_classes'])
