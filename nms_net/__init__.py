"""Drop-in import path: `import nms_net...` resolves to gossipnet_b200.nms_net,
so the reference's drivers (`from nms_net import cfg`, `from nms_net.network
import Gnet`, `from nms_net import matching_module`, `from
nms_net.roi_pooling_layer import roi_pooling_op`) import unchanged.  Every
alias is the very module object of the implementation: one `cfg`, one
parameter scope, whichever name it was imported under."""
import importlib
import sys

_impl = importlib.import_module('gossipnet_b200.nms_net')
for _sub in ('config', 'class_weights', 'matching_module', 'roi_pooling_layer',
             'roi_pooling_layer.roi_pooling_op', 'roi_pooling_layer.roi_pooling_op_grad',
             'network', 'dataset', 'tools'):
    sys.modules['nms_net.' + _sub] = importlib.import_module('gossipnet_b200.nms_net.' + _sub)
sys.modules[__name__] = _impl
