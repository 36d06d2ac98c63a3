"""Where do tensor-map TMA loads put their bytes?  Runs gn_selftest_tma and, instead of
asserting the expected SWIZZLE_128B layout, LOCATES every 16-byte source chunk in the raw
shared-memory dump and prints the mapping it finds (diagnostic for gn_block_tma.cu)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gossipnet_b200 import ops

rng = np.random.RandomState(3)
rows = 1000
mat = torch.from_numpy(rng.uniform(-1, 1, (rows, 64)).astype(np.float32)).cuda().to(torch.bfloat16)
wmat = torch.from_numpy(rng.uniform(-1, 1, (64, 64)).astype(np.float32)).cuda().to(torch.bfloat16)
idx = rng.randint(0, rows, 128).astype(np.int32)
row0 = 256
dump, d = ops.selftest_tma(mat, wmat, torch.from_numpy(idx).cuda(), row0)
torch.cuda.synchronize()
dump = dump.cpu().numpy()
mb = mat.cpu().view(torch.int16).numpy().view(np.uint8).reshape(rows, 8, 16)
wb = wmat.cpu().view(torch.int16).numpy().view(np.uint8).reshape(64, 8, 16)


def locate(region, src_rows, name):
    chunks = region.reshape(-1, 16)
    table = {bytes(c): i for i, c in enumerate(chunks)}
    ok = miss = 0
    bad = []
    for r in range(src_rows.shape[0]):
        for j in range(8):
            pos = table.get(bytes(src_rows[r, j]))
            exp = r * 8 + (j ^ (r & 7))
            if pos is None:
                miss += 1
            elif pos == exp:
                ok += 1
            elif len(bad) < 12:
                bad.append((r, j, pos // 8, pos % 8))
    print('%-8s expected-position %d, elsewhere %d, missing %d of %d' % (
        name, ok, src_rows.shape[0] * 8 - ok - miss, miss, src_rows.shape[0] * 8))
    for r, j, pr, pj in bad:
        print('   src row %d chunk %d -> dump row %d chunk %d' % (r, j, pr, pj))


locate(dump[:16384], mb[row0:row0 + 128], 'tile')
locate(dump[16384:32768], mb[idx], 'gather4')
locate(dump[32768:], wb, 'weights')
a = mat.float().cpu().numpy().astype(np.float64)
ref = (a[row0:row0 + 128] + a[idx]) @ wmat.float().cpu().numpy().astype(np.float64).T
err = np.abs(d.cpu().numpy() - ref)
print('UMMA sw128: max err %.3e (max ref %.3f); rows with err > 1e-3: %d' % (
    err.max(), np.abs(ref).max(), int((err.max(axis=1) > 1e-3).sum())))
