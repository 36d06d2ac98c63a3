"""DetectionMatching kernel alone on the training batch (8 images x N=1000): random scores,
the scores of a fresh network, and tied scores (single-thread introsort fallback)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gossipnet_b200 import ops, synthetic
from gossipnet_b200.nms_net.config import cfg, cfg_from_file
from gossipnet_b200.nms_net.network import Gnet

cfg_from_file(os.path.join(os.path.dirname(__file__), 'coco_person', 'conf.yaml'))
cfg.gnet.num_blocks = 16
imgs = [synthetic.make_image(1000, 1, image_index=i) for i in range(8)]
net = Gnet(1)
res = net.run_batch(imgs)
pred = res['prediction'].reshape(-1)
print('G per image:', [len(im['gt_boxes']) for im in imgs], 'distinct logits', int(torch.unique(pred).numel()), 'of', pred.numel())
eng = net.engine
# rebuild the matching inputs the way engine.match does
import inspect
src = inspect.getsource(type(eng).forward) if hasattr(type(eng), 'forward') else ''
d = lambda a, dt=None: torch.from_numpy(np.ascontiguousarray(a)).cuda() if dt is None else torch.from_numpy(np.ascontiguousarray(a).astype(dt)).cuda()
ious, iou_off, off = [], [0], 0
for im in imgs:
    a = d(im['dets'], np.float32); b = d(np.asarray(im['gt_boxes'], np.float32).reshape(-1, 4))
    m = ops.iou_dense(a, b, crowd=d(np.asarray(im['gt_crowd']).astype(np.uint8)))
    ious.append(m.reshape(-1)); off += m.numel(); iou_off.append(off)
iou = torch.cat(ious)
iou_off = torch.tensor(iou_off, dtype=torch.int64, device='cuda')
ignore = d(np.concatenate([np.asarray(im['gt_crowd']).astype(np.uint8) for im in imgs]))
img_off = torch.tensor(np.arange(9) * 1000, dtype=torch.int32, device='cuda')
gt_off = torch.tensor(np.concatenate([[0], np.cumsum([len(im['gt_boxes']) for im in imgs])]), dtype=torch.int32, device='cuda')
max_gt = max(len(im['gt_boxes']) for im in imgs)
rs = np.random.RandomState(0)
cases = {'random scores': d(rs.rand(8000).astype(np.float32)), 'network logits': pred.contiguous(),
         'tied scores': d((rs.randint(0, 50, 8000) / 50.0).astype(np.float32))}
for name, sc in cases.items():
    ts = []
    for _ in range(8):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ops.detection_matching_batched(iou, iou_off, sc, ignore, img_off, gt_off, max_gt)
        b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    print('%-16s %.1f us' % (name, 1e3 * float(np.median(ts[2:]))))

import ctypes
from gossipnet_b200 import _lib
lib = _lib.load()
if hasattr(lib, 'gn_detection_matching_trace'):
    buf = (ctypes.c_longlong * 8)()
    lib.gn_detection_matching_trace.argtypes = [ctypes.c_void_p]
    ops.detection_matching_batched(iou, iou_off, cases['random scores'], ignore, img_off, gt_off, max_gt)
    torch.cuda.synchronize()
    lib.gn_detection_matching_trace(buf)
    names = ['start', 'init done', 'ranks done', 'gt sort + tie check done', 'records done', 'greedy done']
    for i, nm in enumerate(names):
        print('%9d  %+9d  %s' % (buf[i] - buf[0], buf[i] - buf[i - 1] if i else 0, nm))
