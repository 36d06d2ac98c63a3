"""Programmatic dependent launch on / off: logits identical, graph-replayed forward time."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from gossipnet_b200 import ops
from gossipnet_b200.nms_net.network import Gnet
from gossipnet_b200.session import InferenceSession

bench.setup_cfg(16)
imgs, dets, scores, classes, img_off = bench.make_inputs(64, 1000, 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
res = {}
for rep in range(3):
    for pdl in (False, True):
        ops.set_pdl(pdl)
        net = Gnet(1)
        sess = InferenceSession(net)
        for _ in range(3):
            pred = sess.run(dets, scores, classes, img_off).copy()
        ts = []
        with torch.cuda.stream(sess.stream):
            for _ in range(42):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(sess.stream)
                sess._graph.replay()
                b.record(sess.stream)
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
        print('pdl=%d  forward %.3f ms (median of 40), logit checksum %.6f' % (pdl, float(np.median(ts[2:])), float(pred.astype(np.float64).sum())))
        res[pdl] = pred
print('identical logits:', np.array_equal(res[False], res[True]))
