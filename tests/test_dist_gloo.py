"""CPU, world_size 2 over gloo: the N>1 path's host logic - image sharding and the
single flat gradient all-reduce - reproduces the one-process result.  Each rank
computes the (oracle) gradients of ITS images, the ranks all-reduce gradients +
image count through gossipnet_b200.parallel exactly like Trainer.apply_gradients,
and the mean must equal the full-batch oracle gradient."""
import os
import socket
import sys

import numpy as np
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _image_grad(image_index):
    """flat float64 gradient of one image's loss (oracle), tiny model."""
    from gossipnet_b200 import params as P
    from gossipnet_b200 import synthetic
    from gossipnet_b200.nms_net.config import cfg, reset_cfg
    from oracle import gnet_grad_oracle as gg
    from oracle import gnet_oracle as go
    reset_cfg()
    g = cfg.gnet
    g.num_blocks, g.shortcut_dim, g.reduced_dim, g.pairfeat_dim = 2, 8, 4, 6
    g.num_pwfeat_fc, g.pwfeat_dim, g.pwfeat_narrow_dim, g.predict_fc_dim = 2, 7, 5, 8
    layout, total = P.param_layout(1, cfg)
    flat = P.init_flat(layout, total, cfg, seed=5).astype(np.float64)
    img = synthetic.make_image(12 + image_index, 1, image_index=image_index)
    bd = go.xyxy_to_boxdata(img['dets'])
    m = go.iou(bd, bd)
    pairs = go.neighbor_pairs(m, 0.2)
    raw = go.geometry_feats(bd, m, img['det_scores'], img['det_classes'], pairs, 1, 1.0)
    rs = np.random.RandomState(image_index)
    n = img['dets'].shape[0]
    labels = (rs.uniform(0, 1, n) < 0.4).astype(np.float32)
    weights = np.ones(n, np.float32)
    grads, _ = gg.gradients(P.views(layout, flat), cfg, pairs, raw, n, labels, weights)
    out = np.zeros(total, dtype=np.float64)
    for e in layout.values():
        out[e.offset:e.offset + e.size] = grads[e.name].reshape(-1)
    return out


def _worker(rank, world, port, n_images, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    from gossipnet_b200 import parallel
    r, w, _ = parallel.init_from_env('gloo')
    assert (r, w) == (rank, world) and parallel.world() == world and parallel.rank() == rank
    mine = parallel.shard(list(range(n_images)))
    acc = None
    for i in mine:
        gi = _image_grad(i)
        acc = gi if acc is None else acc + gi
    buf = torch.zeros(acc.shape[0] + 1, dtype=torch.float64)
    buf[:-1] = torch.from_numpy(acc)
    buf[-1] = len(mine)                      # image count rides in the same buffer
    parallel.allreduce_sum_(buf)
    assert parallel.max_over_ranks(rank) == world - 1
    np.save(os.path.join(out_dir, 'rank%d.npy' % rank), buf.numpy())
    torch.distributed.destroy_process_group()


def test_sharded_gradients_allreduce_to_full_batch(tmp_path):
    n_images, world = 5, 2
    mp.spawn(_worker, args=(world, _free_port(), n_images, str(tmp_path)), nprocs=world, join=True)
    bufs = [np.load(str(tmp_path / ('rank%d.npy' % r))) for r in range(world)]
    assert np.array_equal(bufs[0], bufs[1])                 # every rank holds the same sum
    assert bufs[0][-1] == n_images
    full = sum(_image_grad(i) for i in range(n_images))
    assert np.allclose(bufs[0][:-1], full, rtol=1e-12, atol=1e-14)


def test_shard_is_a_balanced_partition():
    from gossipnet_b200 import parallel
    for n in (0, 1, 7, 64):
        for w in (1, 2, 3, 8):
            parts = [parallel.shard(list(range(n)), r, w) for r in range(w)]
            assert sum(parts, []) == list(range(n))
            assert max(map(len, parts)) - min(map(len, parts)) <= 1
