import sys, torch, numpy as np
sys.path.insert(0, '/root/repo')
from gossipnet_b200 import ops, synthetic
for n in (10000, 4000, 1000):
    B = max(1, int(4e8 // (4 * n * n)))
    d = torch.from_numpy(np.stack([synthetic.make_image(n, 1, image_index=i)['dets'] for i in range(min(B, 4))])).cuda()
    d = d.repeat((B + d.shape[0] - 1) // d.shape[0], 1, 1)[:B].contiguous()
    out = torch.empty((B, n, n), device='cuda')
    best = 1e9
    for _ in range(8):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ops.iou_dense(d, d, out=out); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    print('N=%5d B=%3d: %.3f ms %.0f GB/s' % (n, B, best, 4.0 * B * n * n / best / 1e6))
