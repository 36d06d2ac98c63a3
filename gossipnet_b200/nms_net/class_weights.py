"""Per-class loss weights (reference nms_net/class_weights.py:12-23).

The expected weight mass is `1 - pos_weight` for background and `pos_weight`
spread evenly over the foreground classes; index 0 is background.  Counts come
from imdb.tools.get_class_counts like in the reference (or from a precomputed
`class_counts` entry of the imdb dict).
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
    """class_weights.py:12-23."""
    counts = imdb.get('class_counts')
    if counts is None:
        from gossipnet_b200.imdb.tools import get_class_counts
        counts = get_class_counts(imdb)
    return class_equal_weights_from_counts(counts, imdb['num_classes'])
