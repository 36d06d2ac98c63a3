"""Training step time at the per-GPU share of BASELINE configs[3] (64 images x N=1000 over
8 GPUs = 8 images per GPU, 16 blocks): forward with kept activations + matching + loss +
backward + fused Adam on the unfused CUDA pieces."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gossipnet_b200 import synthetic
from gossipnet_b200.nms_net.config import cfg, cfg_from_file
from gossipnet_b200.nms_net.network import Gnet
from gossipnet_b200.trainer import Trainer

cfg_from_file(os.path.join(os.path.dirname(__file__), 'coco_person', 'conf.yaml'))
cfg.gnet.num_blocks = 16
imgs = [synthetic.make_image(1000, 1, image_index=i) for i in range(8)]
net = Gnet(1)
tr = Trainer(net)
for _ in range(2):
    res = tr.step(imgs, 1e-4)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    res = tr.step(imgs, 1e-4)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 3
print('8 images x N=1000, 16 blocks: %.1f ms per training step (%.0f detections/s), P=%d, peak mem %.1f GB'
      % (1e3 * dt, 8000 / dt, res['P'], torch.cuda.max_memory_allocated() / 2**30))
