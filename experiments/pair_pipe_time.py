"""Pair stage of one block on the bench workload: the pipelined kernel
(gn_block_pair_fwd_pipe) against the two-CTA kernel (gn_block_pair_fwd_hl);
both must produce identical pooled features."""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from gossipnet_b200 import _lib, ops
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
image, table, (pair_off, det_off, pair_b, det_b) = eng._operand_images()
wimg = image[pair_off[0]:pair_off[0] + pair_b]
s1 = 'gnet/block1/'
b1, b2 = eng.p[s1 + 'pw_fc1/biases'], eng.p[s1 + 'pw_fc2/biases']
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
out = {}
for mode in ('hl', 'pipe'):
    pooled = torch.zeros((T, 64), device='cuda')
    times = []
    for rep in range(8):
        flush.zero_(); pooled.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        if mode == 'hl':
            ops.block_pair_fwd(res['pw_feats'], red, red, res['pair_c'], res['pair_n'], res['num_pairs'],
                               res['capacity'], None, b1, None, b2, pooled, wimg=wimg)
        else:
            ops.block_pair_fwd_pipe(res['pw_feats'], red, res['pair_c'], res['pair_n'], res['num_pairs'],
                                    res['capacity'], b1, b2, wimg, pooled)
        b.record(); torch.cuda.synchronize()
        times.append(a.elapsed_time(b))
    out[mode] = pooled.clone()
    print('%-5s %.1f us' % (mode, 1e3 * float(np.median(times[2:]))))
print('identical:', bool(torch.equal(out['hl'], out['pipe'])), ' max diff', float((out['hl'] - out['pipe']).abs().max()))
lib = _lib.load()
if hasattr(lib, 'gn_block_pair_trace'):
    buf = (ctypes.c_longlong * 64)()
    lib.gn_block_pair_trace.argtypes = [ctypes.c_void_p]
    lib.gn_block_pair_trace(buf)
    tr = np.array(buf[:], dtype=np.int64)
    names = {2: 'mma a_full (FC1)', 3: 'mma issued FC1', 20: 'fill a_empty', 21: 'fill stored', 0: 'mma h1_full', 1: 'mma issued FC2', 8: 'epi fc1_done', 9: 'epi h1 produced',
             10: 'epi fc2_done', 11: 'epi h2 staged', 12: 'epi pooled'}
    ev = [(tr[i + 32 * par], 'tile%d %s' % (4 + par, nm)) for par in (0, 1) for i, nm in names.items()]
    t0 = min(e[0] for e in ev)
    for tt, nm in sorted(ev):
        print('%7d  %s' % (tt - t0, nm))
