"""One launch of the symmetric dense-IoU kernel per sweep size of bench.py (same batch sizes:
every launch writes >= 256 MB), for an ncu --set full capture of dram bytes per launch."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gossipnet_b200 import ops, synthetic
for n in (1000, 2000, 4000, 8000, 10000):
    b = max(1, int(np.ceil((256 << 20) / (4.0 * n * n))))
    d = torch.from_numpy(np.stack([synthetic.make_image(n, 1, image_index=i)['dets'] for i in range(min(b, 4))])).cuda()
    d = d.repeat((b + d.shape[0] - 1) // d.shape[0], 1, 1)[:b].contiguous()
    out = torch.empty((b, n, n), device='cuda')
    ops.iou_dense(d, d, out=out)
    torch.cuda.synchronize()
    print(n, b)
    del out, d
