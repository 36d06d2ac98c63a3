"""Pair stage on the bench workload: the TMA-fed kernel (gn_block_pair_fwd_tma, pair_mode
'tma') against the round-1 pipeline (gn_block_pair_fwd_pipe): logits of both forwards,
isolated kernel times, whole-forward times; optional clock64 trace (-DBT_TRACE build)."""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from gossipnet_b200 import _lib, ops
from gossipnet_b200.nms_net.network import Gnet

blocks = int(sys.argv[1]) if len(sys.argv) > 1 else 16
images = int(sys.argv[2]) if len(sys.argv) > 2 else 64
bench.setup_cfg(blocks)
imgs, dets, scores, classes, img_off = bench.make_inputs(images, 1000, 0)
net = Gnet(1)
eng = net.engine
d = lambda a: torch.from_numpy(a).cuda()
dd, ds, dc, do = d(dets), d(scores), d(classes), d(img_off)
T = dets.shape[0]
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
preds, fwd_ms = {}, {}
for mode in ('pipe', 'tma'):
    eng.pair_mode = mode
    eng.want_pw_f32 = mode == 'pipe'
    for _ in range(3):
        res = eng.forward(dd, ds, dc, do, max_img=1024)
        try:
            eng.check_overflow()
        except Exception:
            pass
    times = []
    for rep in range(6):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        res = eng.forward(dd, ds, dc, do, max_img=1024)
        b.record(); torch.cuda.synchronize()
        times.append(a.elapsed_time(b))
    fwd_ms[mode] = float(np.median(times[1:]))
    preds[mode] = res['prediction'].clone()
    print('%-5s forward (eager, not graphed) %.3f ms' % (mode, fwd_ms[mode]))
P = int(res['num_pairs'].item())
pa, pb = preds['pipe'].double(), preds['tma'].double()
print('P = %d; logits tma vs pipe: max|d| / max|ref| = %.3e' % (
    P, float((pa - pb).abs().max() / pa.abs().max())))

# isolated kernels on the operands the last forwards left behind
s1 = 'gnet/block1/'
b1, b2 = eng.p[s1 + 'pw_fc1/biases'], eng.p[s1 + 'pw_fc2/biases']
red_all = eng._ws['red_hl'][:(T + 1) * 64].view(T + 1, 64)
ab = eng._ws['u'][:T * 64].view(T, 64)
tma_image, _ = eng._tma_images()
tb = ops.pair_tma_image_bytes()
eng.pair_mode = 'pipe'
image, table, (pair_off, det_off, pair_b, det_b) = eng._operand_images()
ops.prepare_operands(eng.flat, table, image)
wimg = image[pair_off[0]:pair_off[0] + pair_b]
pw = eng._ws['pw'][:res['capacity'] * 32].view(res['capacity'], 32)
for mode in ('pipe', 'tma'):
    pooled = torch.zeros((T, 64), device='cuda')
    times = []
    for rep in range(10):
        flush.zero_(); pooled.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        if mode == 'pipe':
            ops.block_pair_fwd_pipe(pw, red_all[:T], res['pair_c'], res['pair_n'], res['num_pairs'],
                                    res['capacity'], b1, b2, wimg, pooled)
        else:
            ops.block_pair_fwd_tma(eng._pw_hl, red_all, T, ab, res['pair_c'], res['pair_n'],
                                   res['num_pairs'], res['capacity'], b2, tma_image[:tb], pooled)
        b.record(); torch.cuda.synchronize()
        times.append(a.elapsed_time(b))
    us = 1e3 * float(np.median(times[2:]))
    print('%-5s pair kernel %.1f us  (%.1f TFLOP/s algorithmic, %d cycles per 128-pair tile per SM)' % (
        mode, us, 20480.0 * P / us / 1e6, int(us * 1e-6 * 1.965e9 / (P / 128.0 / 148.0))))
lib = _lib.load()
if hasattr(lib, 'gn_block_pair_tma_trace'):
    buf = (ctypes.c_longlong * 64)()
    lib.gn_block_pair_tma_trace.argtypes = [ctypes.c_void_p]
    lib.gn_block_pair_tma_trace(buf)
    tr = np.array(buf[:], dtype=np.int64)
    names = {2: 'mma1 a_full+d1_free', 3: 'mma1 issued FC1', 0: 'mma2 h1_full+d2_free', 1: 'mma2 issued FC2',
             8: 'epi1 fc1_done', 9: 'epi1 h1 produced', 10: 'epi2 fc2_done', 11: 'epi2 staged',
             12: 'pool stg_full', 13: 'pool done'}
    if hasattr(lib, 'gn_block_pair_tma_acc'):
        ab_ = (ctypes.c_longlong * 32)()
        lib.gn_block_pair_tma_acc.argtypes = [ctypes.c_void_p]
        lib.gn_block_pair_tma_acc(ab_)
        acc = np.array(ab_[:], dtype=np.int64)
        ntile = (P + 127) // 128 / 148.0
        lab = {0: 'producer a_empty wait', 1: 'producer cp.async wait', 4: 'mma1 a_full wait', 5: 'mma1 d1_free wait', 8: 'mma2 h1_full wait',
               9: 'mma2 d2_free wait', 12: 'epi1 fc1_done wait', 13: 'epi1 h1-free wait', 16: 'epi2 fc2_done wait',
               17: 'epi2 stg_free wait', 20: 'pool stg_full wait'}
        for k in sorted(lab):
            print('  %-24s %7.0f cycles per tile' % (lab[k], acc[k] / ntile))
    if tr[22] > 0:
        cyc, ns = tr[22] - tr[20], tr[23] - tr[21]
        print('CTA 0, tiles 6..106: %.0f cycles per tile, %.0f ns per tile -> SM clock %.0f MHz' % (
            cyc / 100.0, ns / 100.0, 1e3 * cyc / ns))
    ev = [(tr[i + 32 * par], 'tile%d %s' % (6 + par, nm)) for par in (0, 1) for i, nm in names.items()]
    t0 = min(e[0] for e in ev)
    for tt, nm in sorted(ev):
        print('%7d  %s' % (tt - t0, nm))
