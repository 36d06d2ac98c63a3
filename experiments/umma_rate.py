import torch, sys
sys.path.insert(0,'/root/repo')
from gossipnet_b200 import _lib
out = torch.zeros(2, dtype=torch.int64, device='cuda')
s = torch.cuda.current_stream().cuda_stream
for ctas in (1, 148):
    for n in (32, 64, 128, 256):
        for db in (1, 4):
            _lib.call("gn_selftest_umma_rate", n, 2049, db, ctas, out.data_ptr(), s)
            torch.cuda.synchronize()
            c, r = out.tolist()
            print('ctas=%3d N=%3d distinct_b=%d: %.1f cycles/MMA (floor %d)' % (ctas, n, db, c / r, max(128 * n // 256, 1)))
