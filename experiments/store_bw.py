"""Pure-store bandwidth of one B200 with hand-written kernels (gn_selftest_store_bw) next to
cudaMemset, and the dense IoU kernels at the same sizes: is the IoU kernel at the hardware's
store ceiling?"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gossipnet_b200 import ops, synthetic


def t_ms(fn, reps=6, inner=1):
    out = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(inner):
            fn()
        b.record(); torch.cuda.synchronize()
        out.append(a.elapsed_time(b) / inner)
    return min(out[1:])


for size in (400 << 20, 1 << 30, 2 << 30):
    buf = torch.empty(size, dtype=torch.uint8, device='cuda')
    row = ['%5d MiB' % (size >> 20)]
    row.append('memset %.0f' % (size / t_ms(lambda: buf.zero_()) / 1e6))
    for mode in ops.STORE_BW_MODES:
        best = max((size / t_ms(lambda: ops.selftest_store_bw(buf, mode, c)) / 1e6, c)
                   for c in (2, 4, 8, 16))
        row.append('%s %.0f (x%d)' % (mode, best[0], best[1]))
    print(' | '.join(row), flush=True)
    del buf
for n in (10000,):
    d = torch.from_numpy(synthetic.make_image(n, 1)['dets']).cuda()
    out = torch.empty((1, n, n), device='cuda')
    nbytes = 4.0 * n * n
    du = d.unsqueeze(0)
    print('iou symmetric N=%d: %.0f GB/s (one launch per event pair), %.0f GB/s (8 launches per pair)' % (
        n, nbytes / t_ms(lambda: ops.iou_dense(du, du, out=out)) / 1e6,
        nbytes / t_ms(lambda: ops.iou_dense(du, du, out=out), inner=8) / 1e6))
    d2 = d.clone()
    print('iou general   N=%d: %.0f GB/s' % (n, nbytes / t_ms(lambda: ops.iou_dense(d.unsqueeze(0), d2.unsqueeze(0), out=out)) / 1e6))
