"""Mirror of the reference's `nms_net` package (nms_net/__init__.py:1)."""
from gossipnet_b200.nms_net.config import cfg  # noqa: F401
