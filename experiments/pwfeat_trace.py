"""Phase timeline of one tile of the pipelined pair-feature kernel (CTA 0, third
tile).  Needs the library built with EXTRA=-DPP_TRACE:
    make -C gossipnet_b200/csrc clean && make -C gossipnet_b200/csrc EXTRA=-DPP_TRACE -j8
"""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from gossipnet_b200 import _lib
from gossipnet_b200.nms_net.network import Gnet

bench.setup_cfg(16)
imgs, dets, scores, classes, img_off = bench.make_inputs(64, 1000, 0)
net = Gnet(1)
eng = net.engine
d = lambda a: torch.from_numpy(a).cuda()
dd, ds, dc, do = d(dets), d(scores), d(classes), d(img_off)
eng.neighbors(dd, do)
row_ptr, num_pairs, pair_c, pair_n, pair_iou, cap = eng.neighbors(dd, do)
for _ in range(3):
    eng.pair_features(dd, ds, dc, pair_c, pair_n, pair_iou, num_pairs, cap)
torch.cuda.synchronize()
lib = _lib.load()
buf = (ctypes.c_longlong * 64)()
lib.gn_pwfeat_trace.argtypes = [ctypes.c_void_p]
assert lib.gn_pwfeat_trace(buf) == 0
tr = np.array(buf[:], dtype=np.int64)
t0 = tr[0]
names = {0: 'mma tile start'}
for q in range(4):
    names[1 + 2 * q] = 'mma h_full L2 q%d' % q
    names[2 + 2 * q] = 'mma issued L2 q%d' % q
    names[10 + 2 * q] = 'mma h_full L3 q%d' % q
    names[32 + 2 * q] = 'epi l1_done q%d' % q
    names[33 + 2 * q] = 'epi produced e1 q%d' % q
    names[41 + q] = 'epi produced e2 q%d' % q
names[20] = 'mma committed acc3'
names[40] = 'epi acc2_done'
names[45] = 'epi acc3_done'
names[46] = 'epi output written'
names[47] = 'epi acc3 loaded'
for i in sorted(names, key=lambda i: tr[i]):
    print('%7d  %s' % (tr[i] - t0, names[i]))
