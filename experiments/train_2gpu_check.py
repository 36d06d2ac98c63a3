"""Two-rank NCCL check of the training step: both ranks train on their
shard of every step's images with ONE all-reduce of the flat gradient; parameters must stay
bit-identical across ranks and equal (to fp32 summation order) a single-rank run on all
images.   torchrun --nproc-per-node 2 experiments/train_2gpu_check.py"""
import os, sys
import numpy as np, torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gossipnet_b200 import parallel, synthetic
from gossipnet_b200.nms_net.config import cfg, cfg_from_file
from gossipnet_b200.nms_net.network import Gnet
from gossipnet_b200.trainer import Trainer

rank, world, local = parallel.init_from_env()
cfg_from_file(os.path.join(os.path.dirname(__file__), 'coco_person', 'conf.yaml'))
cfg.gnet.num_blocks = 4
imgs = [synthetic.make_image(300 + 20 * i, 1, image_index=i) for i in range(8)]
net = Gnet(1)
tr = Trainer(net)
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
grad1 = None
for step in range(5):
    if step == 2:
        torch.cuda.synchronize(); t0.record()
    res = tr.step(parallel.shard(imgs), 1e-3)
    if step == 0:
        grad1 = tr.grad.clone()          # the all-reduced gradient of the first step
t1.record(); torch.cuda.synchronize()
flat = net.engine.flat.clone()
gathered = [torch.empty_like(flat) for _ in range(world)]
dist.all_gather(gathered, flat)
same = all(torch.equal(gathered[0], g) for g in gathered)
if rank == 0:
    print('ranks', world, 'images/step', int(res['images_in_step']), 'params identical across ranks:', same,
          ' ms/step %.2f' % (t0.elapsed_time(t1) / 3))
    # single-process reference on all 8 images (same init seed -> same start)
    from gossipnet_b200 import params as P
    dist.barrier()
else:
    dist.barrier()
dist.destroy_process_group()
if rank == 0:
    os.environ['WORLD_SIZE'] = '1'
    net1 = Gnet(1)
    tr1 = Trainer(net1)
    g1 = None
    for step in range(5):
        tr1.step(imgs, 1e-3)
        if step == 0:
            g1 = tr1.grad.clone()
    # the summed gradient itself: equal up to the fp32 summation order (atomics, sharding)
    rel = float((g1 - grad1).norm() / g1.norm())
    print('first-step gradient, single rank vs all-reduced two ranks: relative l2 difference %.3e' % rel)
    # Adam's first steps move every element by ~lr whatever the gradient's size, so a gradient
    # element whose sign depends on the summation order shows up as a difference of ~lr per step
    d = float((net1.engine.flat - flat).abs().max())
    print('max |single-rank - two-rank| parameter after 5 Adam steps (lr 1e-3): %.3e' % d)
    assert same and rel < 1e-5 and d < 5 * 2e-3
    print('OK')
