"""Times gn_pwfeat_mlp_fwd (fp32 semantics and plain bf16) on the bench workload and checks
it against the staged CUDA pieces (pair_geometry + 3 x fc_fwd)."""
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
eng.neighbors(dd, do, 1000)
row_ptr, num_pairs, pair_c, pair_n, pair_iou, cap = eng.neighbors(dd, do, 1000)
P = int(num_pairs.item())
eng.use_fused = False
ref = eng.pair_features(dd, ds, dc, pair_c, pair_n, pair_iou, num_pairs, cap)[:P].clone()
eng.use_fused = True
for bf16 in (False, True):
    eng.bf16 = bf16
    times = []
    for rep in range(6):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        pw = eng.pair_features(dd, ds, dc, pair_c, pair_n, pair_iou, num_pairs, cap)
        b.record(); torch.cuda.synchronize()
        times.append(a.elapsed_time(b))
    err = float((pw[:P] - ref).abs().max() / ref.abs().max())
    print('%s: %.1f us (P=%d)  max rel err vs staged fp32 FCs %.2e' % (
        'bf16 ' if bf16 else 'fp32 ', 1e3 * float(np.median(times[2:])), P, err))
