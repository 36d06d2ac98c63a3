"""GPU: pair features (A4, A5), block pieces (A7) and the FC layer vs the oracle."""
import numpy as np
import pytest
import torch

from gossipnet_b200 import ops, synthetic
from gossipnet_b200 import params as P
from gossipnet_b200.nms_net.config import cfg
from oracle import gnet_oracle as go
from tests.helpers import load_experiment, rel_err

pytestmark = pytest.mark.gpu
F32 = np.float32


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def graph(img):
    d = img['dets']
    bd = go.xyxy_to_boxdata(d)
    m = go.iou(bd, bd)
    pairs = go.neighbor_pairs(m, 0.2)
    return bd, m, pairs


@pytest.mark.parametrize('n,C', [(40, 1), (300, 1), (150, 80), (1000, 1)])
def test_pair_geometry(n, C):
    img = synthetic.make_image(n, C, image_index=1)
    bd, m, pairs = graph(img)
    ref = go.geometry_feats(bd, m, img['det_scores'], img['det_classes'], pairs, C, 1.0)
    Pn = pairs.shape[0]
    got = ops.pair_geometry(dev(img['dets']), dev(img['det_scores']), dev(img['det_classes']),
                            dev(pairs[:, 0], torch.int32), dev(pairs[:, 1], torch.int32),
                            dev(m[pairs[:, 0], pairs[:, 1]]),
                            torch.tensor([Pn], dtype=torch.int32, device='cuda'), Pn, C,
                            1.0).cpu().numpy()
    assert got.shape == ref.shape
    # scores, iou, distances: exactly rounded ops -> bit exact; the three log
    # features go through logf (<= 1 ulp vs numpy's log) then a division
    exact = list(range(0, got.shape[1] - 3))
    assert np.array_equal(got[:, exact], ref[:, exact])
    assert np.allclose(got[:, -3:], ref[:, -3:], rtol=0, atol=2e-6)


@pytest.mark.parametrize('rows,k,n,relu,res', [(1, 9, 256, True, False), (1000, 128, 32, True, False),
                                                (777, 64, 128, True, True), (64, 128, 1, False, False),
                                                (300, 167, 256, True, False), (5, 3, 5, False, True)])
def test_fc_layer(rows, k, n, relu, res):
    rs = np.random.RandomState(rows + k)
    x = rs.normal(0, 1, (rows, k)).astype(F32)
    w = rs.normal(0, 0.2, (k, n)).astype(F32)
    b = rs.normal(0, 0.1, n).astype(F32)
    r = rs.normal(0, 1, (rows, n)).astype(F32) if res else None
    ref = x.astype(np.float64) @ w.astype(np.float64) + b
    if res:
        ref = ref + r
    if relu:
        ref = np.maximum(ref, 0)
    got = ops.fc_fwd(dev(x), dev(w), dev(b), relu, residual=dev(r) if res else None).cpu().numpy()
    assert np.allclose(got, ref, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize('n,C,exp', [(300, 1, 'coco_person'), (150, 80, 'coco_multiclass'),
                                     (1000, 1, 'coco_person')])
@pytest.mark.parametrize('ffma', [False, True])
def test_fused_pair_feature_mlp(n, C, exp, ffma):
    load_experiment(exp)
    layout, total = P.param_layout(C, cfg)
    flat = P.init_flat(layout, total, cfg, seed=3)
    p = P.views(layout, flat)
    img = synthetic.make_image(n, C, image_index=2)
    bd, m, pairs = graph(img)
    raw = go.geometry_feats(bd, m, img['det_scores'], img['det_classes'], pairs, C, 1.0)
    ref = go.pw_feats_fc(raw, p, cfg)
    Pn = pairs.shape[0]
    cap = Pn + 100
    dp = P.views(layout, dev(flat))
    pad = lambda a: np.concatenate([a, np.zeros(100, a.dtype)])
    got = ops.pwfeat_mlp_fwd(
        dev(img['dets']), dev(img['det_scores']), dev(img['det_classes']) if C > 1 else None,
        dev(pad(pairs[:, 0]), torch.int32), dev(pad(pairs[:, 1]), torch.int32),
        dev(pad(m[pairs[:, 0], pairs[:, 1]])), torch.tensor([Pn], dtype=torch.int32, device='cuda'),
        cap, C, 1.0, *[dp['gnet/pw_feats/fc%d/%s' % (i, k)] for i in (1, 2, 3)
                       for k in ('weights', 'biases')],
        out=torch.full((cap, 32), -1.0, device='cuda'), ffma=ffma).cpu().numpy()
    assert np.all(got[Pn:] == -1.0)          # rows past P untouched
    assert rel_err(got[:Pn], ref) < (2e-5 if ffma else 5e-5)   # bf16x3 on the tensor cores


@pytest.mark.parametrize('n', [40, 300, 1000])
def test_block_pair_stage_fused_equals_unfused_equals_oracle(n):
    load_experiment('coco_person')
    rs = np.random.RandomState(n)
    img = synthetic.make_image(n, 1, image_index=4)
    _, _, pairs = graph(img)
    Pn = pairs.shape[0]
    pw = np.maximum(rs.normal(0, 1, (Pn, 32)), 0).astype(F32)
    feats = np.maximum(rs.normal(0, 1, (n, 32)), 0).astype(F32)
    w1 = rs.normal(0, 0.15, (96, 64)).astype(F32)
    b1 = rs.normal(0, 0.1, 64).astype(F32)
    w2 = rs.normal(0, 0.15, (64, 64)).astype(F32)
    b2 = rs.normal(0, 0.1, 64).astype(F32)
    pc, pn = pairs[:, 0], pairs[:, 1]
    nf = feats[pn].copy()
    nf[pc == pn] = 0
    x = np.concatenate([pw, feats[pc], nf], 1)
    h = np.maximum(x.astype(np.float64) @ w1 + b1, 0)
    h = np.maximum(h @ w2.astype(np.float64) + b2, 0)
    ref = go.segment_max(h.astype(F32), pc, n)

    dpc, dpn = dev(pc, torch.int32), dev(pn, torch.int32)
    npairs = torch.tensor([Pn], dtype=torch.int32, device='cuda')
    dfe, dpw = dev(feats), dev(pw)
    pooled = torch.zeros((n, 64), device='cuda')
    ops.block_pair_fwd(dpw, dfe, dfe, dpc, dpn, npairs, Pn, dev(w1), dev(b1), dev(w2), dev(b2),
                       pooled)
    fused = pooled.cpu().numpy()          # tensor cores, bf16x3
    pooled = torch.zeros((n, 64), device='cuda')
    ops.block_pair_fwd(dpw, dfe, dfe, dpc, dpn, npairs, Pn, dev(w1), dev(b1), dev(w2), dev(b2),
                       pooled, ffma=True)
    assert rel_err(pooled.cpu().numpy(), ref) < 1e-5   # fp32 CUDA-core variant

    xg = ops.block_gather_concat(dpw, dfe, dfe, dpc, dpn, npairs, Pn)
    assert np.array_equal(xg.cpu().numpy(), x)
    h1 = ops.fc_fwd(xg, dev(w1), dev(b1), True)
    h2 = ops.fc_fwd(h1, dev(w2), dev(b2), True)
    row_ptr = ops.exclusive_scan(dev(np.bincount(pc, minlength=n), torch.int32))
    unfused = ops.segment_max(h2, row_ptr, n).cpu().numpy()
    assert rel_err(fused, ref) < 3e-5
    assert rel_err(unfused, ref) < 1e-5


def test_segment_max_exact():
    rs = np.random.RandomState(0)
    deg = rs.randint(1, 40, 200)
    ids = np.repeat(np.arange(200), deg)
    x = rs.normal(0, 1, (ids.size, 64)).astype(F32)
    ref = go.segment_max(x, ids, 200)
    row_ptr = ops.exclusive_scan(dev(deg, torch.int32))
    got = ops.segment_max(dev(x), row_ptr, 200).cpu().numpy()
    assert np.array_equal(got, ref)


def test_unsupported_fused_shape_is_reported():
    z = torch.zeros((8, 16), device='cuda')
    i = torch.zeros(8, dtype=torch.int32, device='cuda')
    with pytest.raises(NotImplementedError):
        ops.block_pair_fwd(z, z, z, i, i, i[:1], 8, torch.zeros((48, 64), device='cuda'),
                           torch.zeros(64, device='cuda'), torch.zeros((64, 64), device='cuda'),
                           torch.zeros(64, device='cuda'), torch.zeros((8, 64), device='cuda'))


@pytest.mark.parametrize('n', [1, 100, 128, 1000, 3001])
def test_fused_det_level_kernel(n):
    """gn_block_det_fwd (fc1, fc2, shortcut, next reduce_dim on the tensor cores) vs float64."""
    rs = np.random.RandomState(n)
    pooled = np.maximum(rs.normal(0, 1, (n, 64)), 0).astype(F32)
    feats = np.maximum(rs.normal(0, 1, (n, 128)), 0).astype(F32)
    mk = lambda k, m: (rs.normal(0, 0.15, (k, m)).astype(F32), rs.normal(0, 0.1, m).astype(F32))
    fc1, fc2, rd = mk(64, 64), mk(64, 128), mk(128, 32)
    d1 = np.maximum(pooled.astype(np.float64) @ fc1[0] + fc1[1], 0)
    out = np.maximum(feats + d1 @ fc2[0] + fc2[1], 0)
    red = np.maximum(out @ rd[0] + rd[1], 0)
    dv = lambda pair: (dev(pair[0]), dev(pair[1]))
    d_pooled = dev(pooled)
    feats_out = torch.empty((n, 128), device='cuda')
    red_f32 = torch.empty((n, 32), device='cuda')
    red_hl = torch.empty((n, 64), dtype=torch.bfloat16, device='cuda')
    ops.block_det_fwd(d_pooled, dev(feats), dv(fc1), dv(fc2), dv(rd), feats_out=feats_out,
                      red_f32=red_f32, red_hl=red_hl)
    assert torch.all(d_pooled == 0)                      # re-armed for the next atomicMax pass
    assert rel_err(feats_out.cpu().numpy(), out) < 3e-5
    assert rel_err(red_f32.cpu().numpy(), red) < 3e-5
    hl = red_hl.float().cpu().numpy()
    assert rel_err(hl[:, :32] + hl[:, 32:], red) < 3e-5  # hi + lo reconstructs the fp32 value
    # stage B alone (block 1) and stage A alone (last block)
    red2 = torch.empty((n, 32), device='cuda')
    ops.block_det_fwd(None, dev(feats), None, None, dv(rd), red_f32=red2)
    ref2 = np.maximum(feats.astype(np.float64) @ rd[0] + rd[1], 0)
    assert rel_err(red2.cpu().numpy(), ref2) < 3e-5
    out3 = torch.empty((n, 128), device='cuda')
    ops.block_det_fwd(dev(pooled), dev(feats), dv(fc1), dv(fc2), None, feats_out=out3)
    assert rel_err(out3.cpu().numpy(), out) < 3e-5


@pytest.mark.parametrize('bf16', [False, True])
@pytest.mark.parametrize('n', [1, 127, 129, 1000, 20011])
def test_det_copy_engine_kernel_bit_identical(n, bf16):
    """gn_block_det_fwd_tma (tile transfers by tensor-map TMA from a dedicated warp, pooled
    rows prefetched; 20011 rows = more tiles than SMs, so the per-CTA tile pipeline runs)
    against gn_block_det_fwd_img_u: feats_out, red_hl, u and the re-zeroed pooled buffer,
    with and without stage B."""
    from gossipnet_b200.nms_net.network import Gnet
    load_experiment('coco_person', num_blocks=2)
    eng = Gnet(1).engine
    image, _, (_, det_off, _, det_b) = eng._operand_images()
    eng._prepare_block_images()
    p = eng.p
    rs = np.random.RandomState(n)
    pooled0 = dev(np.maximum(rs.normal(0, 1, (n, 64)), 0).astype(F32))
    feats = dev(np.maximum(rs.normal(0, 1, (n, 128)), 0).astype(F32))
    for b, last in ((0, False), (1, False), (2, True)):      # block 1's reduce only / both stages / last block
        wimg = image[det_off[b]:det_off[b] + det_b]
        s, nxt = 'gnet/block%d/' % b, 'gnet/block%d/' % (b + 1)
        got = []
        for fn in (ops.block_det_fwd_img_u, ops.block_det_fwd_tma):
            pooled = pooled0.clone() if b >= 1 else None
            out = torch.full((n, 128), -1.0, device='cuda')
            red = torch.full((n + 1, 64), -1.0, device='cuda').to(torch.bfloat16)
            u = torch.full((n, 64), -1.0, device='cuda')
            fn(pooled, feats, wimg, p[s + 'fc1/biases'] if b >= 1 else None,
               p[s + 'fc2/biases'] if b >= 1 else None,
               None if last else p[nxt + 'reduce_dim/biases'], out if b >= 1 else None,
               None if last else red[:n], None if last else p[nxt + 'pw_fc1/biases'],
               None if last else u, bf16=bf16)
            if b >= 1:
                assert torch.all(pooled == 0)
            got.append((out, red, u))
        for a, c in zip(*got):
            assert torch.equal(a, c)
        assert float(got[1][0 if b >= 1 else 2].abs().sum()) > 0
        assert torch.all(got[1][1][n] == -1.0)        # the row behind the last detection is not touched
