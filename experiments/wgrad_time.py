"""gn_fc_bwd_weight_tc on the training shapes (P = 311 072 rows): time per call and
effective HBM bandwidth."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gossipnet_b200 import ops
rows = 311072
for k, n in ((64, 64), (96, 64), (256, 32), (256, 256), (9, 256)):
    x = torch.randn(rows, k, device='cuda'); dy = torch.randn(rows, n, device='cuda')
    y = torch.randn(rows, n, device='cuda')
    dw = torch.zeros(k, n, device='cuda'); db = torch.zeros(n, device='cuda')
    ts = []
    for rep in range(6):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ops.fc_bwd_weight_tc(x, dy, dw, db, mask=y); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = min(ts[1:])
    print('wgrad k=%3d n=%3d: %.1f us, %.2f TB/s' % (k, n, 1e3 * ms, rows * (k + 2 * n) * 4 / ms / 1e9))
