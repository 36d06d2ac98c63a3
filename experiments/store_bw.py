"""Pure-store bandwidth reference for the dense IoU kernel (write-only, 400 MB)."""
import torch
x = torch.empty(100_000_000, dtype=torch.float32, device='cuda')
y = torch.empty_like(x)
def t(fn, n=10):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best
ms = t(lambda: x.zero_()); print('zero_ 400MB: %.3f ms  %.0f GB/s (write only)' % (ms, 0.4 / ms * 1e3))
ms = t(lambda: x.fill_(1.5)); print('fill_ 400MB: %.3f ms  %.0f GB/s (write only)' % (ms, 0.4 / ms * 1e3))
ms = t(lambda: y.copy_(x)); print('copy 400MB: %.3f ms  %.0f GB/s (read+write)' % (ms, 0.8 / ms * 1e3))
big = torch.empty(1 << 30, dtype=torch.bfloat16, device='cuda'); big2 = torch.empty_like(big)
ms = t(lambda: big2.copy_(big)); print('copy 2GiB: %.3f ms  %.0f GB/s (read+write)' % (ms, 4.295 / ms * 1e3))
ms = t(lambda: big.zero_()); print('zero_ 2GiB: %.3f ms  %.0f GB/s (write only)' % (ms, 2.147 / ms * 1e3))
