"""Detection-level kernel of the bench workload timed alone (as bench.kernel_rooflines does),
plus the per-phase clock64 trace of one tile when the library was built with -DDT_TRACE."""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from gossipnet_b200 import _lib, ops
from gossipnet_b200.nms_net.network import Gnet
from gossipnet_b200.session import InferenceSession

images = int(sys.argv[1]) if len(sys.argv) > 1 else 64
bench.setup_cfg(16)
imgs, dets, scores, classes, img_off = bench.make_inputs(images, 1000, 0)
net = Gnet(1)
sess = InferenceSession(net, use_graph=False)
for _ in range(2):
    sess.run(dets, scores, classes, img_off)
P, T = int(sess.h_np[0]), dets.shape[0]
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
peaks = bench._load_json('MEASURED_PEAKS.json')
with torch.cuda.stream(sess.stream):
    r = bench.kernel_rooflines(torch, ops, net.engine, sess, sess.stream, flush, P, T, 16, False,
                               peaks, 'measured')
for k in ('roofline', 'roofline_pwfeat', 'roofline_det'):
    print('%-16s %-24s %8.1f us  frac %.3f' % (k, r[k]['kernel'], 1e3 * r[k]['ms_per_launch'], r[k]['frac']))
lib = _lib.load()


def det_warm():
    """the det launch again without the L2 flush (inputs L2 resident, as inside the step)"""
    eng = net.engine
    image, _, (_, det_off, _, det_b) = eng._operand_images()
    p = eng.p
    pooled = eng._buf('pooled', (T, 64))
    feats, outb = eng._buf('feats0', (T, 128)), eng._buf('feats1', (T, 128))
    inter = eng._ws['red_hl'][:T * 64].view(T, 64)
    u = eng._buf('u', (T, 64))
    fn = ops.block_det_fwd_tma if eng.det_tma else ops.block_det_fwd_img_u
    ts = []
    for _ in range(8):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn(pooled, feats, image[det_off[1]:det_off[1] + det_b], p['gnet/block1/fc1/biases'],
           p['gnet/block1/fc2/biases'], p['gnet/block2/reduce_dim/biases'], outb, inter,
           p['gnet/block2/pw_fc1/biases'], u)
        b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    print('det kernel, L2 warm: %.1f us' % (1e3 * float(np.median(ts[2:]))))


with torch.cuda.stream(sess.stream):
    det_warm()
if hasattr(lib, 'gn_block_det_tma_trace'):
    buf = (ctypes.c_longlong * 32)()
    lib.gn_block_det_tma_trace.argtypes = [ctypes.c_void_p]
    lib.gn_block_det_tma_trace(buf)
    names = ['tile start', 'pooled split -> A, sync', 'fc1 done', 'epilogue 1, sync', 'shortcut + U boxes free',
             'fc2 done', 'epilogue 2, sync', 'reduce_dim done', 'epilogue red, sync', 'U gemm done', 'U staged, sync']
    t0 = buf[0]
    for i, n in enumerate(names):
        print('%6d  %+6d  %s' % (buf[i] - t0, buf[i] - buf[i - 1] if i else 0, n))
    print('head: split stored %d, fence done %d' % (buf[11] - t0, buf[12] - t0))
if hasattr(lib, 'gn_block_det_trace'):
    buf = (ctypes.c_longlong * 32)()
    lib.gn_block_det_trace.argtypes = [ctypes.c_void_p]
    lib.gn_block_det_trace(buf)
    names = ['tile start', 'pooled tile loaded + split', 'sync', 'fc1 done', 'epilogue 1 (d1 -> A)', 'sync',
             'shortcut rows staged', 'sync', 'fc2 done', 'epilogue 2 (feats_out)', 'sync', 'feats tile stored',
             'reduce_dim done', 'epilogue red', 'sync', 'U gemm done', 'U stored', 'sync (tile end)']
    t0 = buf[0]
    for i, n in enumerate(names):
        print('%6d  %+6d  %s' % (buf[i] - t0, buf[i] - buf[i - 1] if i else 0, n))
