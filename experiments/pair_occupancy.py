"""How much does the second resident CTA buy the pair kernel?  Times
gn_block_pair_fwd_hl on the bench workload with 1 and 2 CTAs per SM
(GN_PAIR_CTAS_PER_SM, experiments only)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from gossipnet_b200 import ops
from gossipnet_b200.nms_net.network import Gnet

bench.setup_cfg(16)
imgs, dets, scores, classes, img_off = bench.make_inputs(64, 1000, 0)
net = Gnet(1)
eng = net.engine
d = lambda a: torch.from_numpy(a).cuda()
dd, ds, dc, do = d(dets), d(scores), d(classes), d(img_off)
for _ in range(2):
    res = eng.forward(dd, ds, dc, do)
    try:
        eng.check_overflow()
    except Exception:
        pass
res = eng.forward(dd, ds, dc, do)
T = dets.shape[0]
red = eng._ws['red_hl'][:T * 64].view(T, 64)
pooled = eng._buf('pooled', (T, 64))
s1 = 'gnet/block1/'
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
for per_sm in (1, 2):
    os.environ['GN_PAIR_CTAS_PER_SM'] = str(per_sm)
    times = []
    for rep in range(8):
        flush.zero_(); pooled.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ops.block_pair_fwd(res['pw_feats'], red, red, res['pair_c'], res['pair_n'], res['num_pairs'],
                           res['capacity'], eng.p[s1 + 'pw_fc1/weights'], eng.p[s1 + 'pw_fc1/biases'],
                           eng.p[s1 + 'pw_fc2/weights'], eng.p[s1 + 'pw_fc2/biases'], pooled)
        b.record(); torch.cuda.synchronize()
        times.append(a.elapsed_time(b))
    print('CTAs/SM = %d: %.1f us' % (per_sm, 1e3 * float(np.median(times[2:]))))
