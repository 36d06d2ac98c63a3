"""Dense det x det IoU at N = 10 000 (one image): a few launches of the symmetric kernel
and of the general kernel, for ncu captures / quick timing."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gossipnet_b200 import ops, synthetic
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
d = torch.from_numpy(synthetic.make_image(n, 1)['dets']).cuda().unsqueeze(0)
d2 = d.clone()
out = torch.empty((1, n, n), device='cuda')
for _ in range(3):
    ops.iou_dense(d, d, out=out)
    ops.iou_dense(d, d2, out=out)
torch.cuda.synchronize()
