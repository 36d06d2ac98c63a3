"""Times gn_pwfeat_mlp_fwd on the bench workload: pipelined kernel vs the
unpipelined one (GN_PWFEAT_V1=1), and checks they agree."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from gossipnet_b200.nms_net.network import Gnet

bench.setup_cfg(16)
imgs, dets, scores, classes, img_off = bench.make_inputs(64, 1000, 0)
net = Gnet(1)
eng = net.engine
d = lambda a: torch.from_numpy(a).cuda()
dd, ds, dc, do = d(dets), d(scores), d(classes), d(img_off)
row_ptr, num_pairs, pair_c, pair_n, pair_iou, cap = eng.neighbors(dd, do)
eng.capacity = 0
row_ptr, num_pairs, pair_c, pair_n, pair_iou, cap = eng.neighbors(dd, do)
P = int(num_pairs.item())
outs = {}
for v1 in ('1', '0'):
    os.environ['GN_PWFEAT_V1'] = v1
    times = []
    for rep in range(6):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        pw = eng.pair_features(dd, ds, dc, pair_c, pair_n, pair_iou, num_pairs, cap)
        b.record(); torch.cuda.synchronize()
        times.append(a.elapsed_time(b))
    outs[v1] = pw[:P].clone()
    print('GN_PWFEAT_V1=%s: %.1f us (P=%d)' % (v1, 1e3 * float(np.median(times[2:])), P))
print('max |v2 - v1| =', float((outs['0'] - outs['1']).abs().max()), ' max |v1| =', float(outs['1'].abs().max()))
