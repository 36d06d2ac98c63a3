"""ORACLE SUPPORT: placeholder for slim.nets (resnet_v1 is never called)."""
from tensorflow.contrib.slim.nets import resnet_v1  # noqa: F401
