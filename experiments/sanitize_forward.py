"""Small workloads that touch every kernel of the path, for compute-sanitizer:
    compute-sanitizer --tool memcheck python experiments/sanitize_forward.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gossipnet_b200 import synthetic
from gossipnet_b200.nms_net.config import cfg, cfg_from_file, reset_cfg
from gossipnet_b200.nms_net.network import Gnet
from gossipnet_b200.session import InferenceSession
from gossipnet_b200.trainer import Trainer

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def setup(exp, blocks, **kw):
    reset_cfg()
    cfg_from_file(os.path.join(ROOT, 'experiments', exp, 'conf.yaml'))
    cfg.gnet.num_blocks = blocks
    for k, v in kw.items():
        cfg.gnet[k] = v


# ragged batch, fp32 semantics, with ground truth (matching + loss)
setup('coco_person', 3)
net = Gnet(1)
imgs = [synthetic.make_image(n, 1, image_index=i) for i, n in enumerate((300, 1, 57, 640))]
res = net.run_batch(imgs)
print('ragged fp32   ', float(res['prediction'].sum()), res['P'])
# host-buffer session (CUDA graphs)
sess = InferenceSession(net)
d = np.concatenate([im['dets'] for im in imgs]); s = np.concatenate([im['det_scores'] for im in imgs])
c = np.concatenate([im['det_classes'] for im in imgs]); off = np.array([0, 300, 301, 358, 998], np.int32)
print('session       ', float(sess.run(d, s, c, off).sum()))
# pipelined session: copy streams + captured forwards (two slots), three batches
outs = [o.copy() for o in sess.run_pipelined([(d, s, c, off)] * 3)]
print('pipelined     ', float(outs[-1].sum()), bool(np.array_equal(outs[0], outs[2])))
# plain bf16
setup('coco_person', 3, compute_dtype='bf16')
print('bf16          ', float(Gnet(1)(imgs[0]).sum()))
# multi-class
setup('coco_multiclass', 2)
print('multiclass    ', float(Gnet(80)(synthetic.make_image(200, 80)).sum()))
# image features + a training step
setup('coco_person', 2, imfeats=True, imfeat_channels=8, imfeat_dim=16)
net = Gnet(1)
img = synthetic.make_image(150, 1)
img['imfeats'] = np.random.RandomState(0).normal(size=(1, 38, 63, 8)).astype(np.float32)
tr = Trainer(net)
out = tr.step([img], 1e-3)
print('imfeats train ', float(out['loss_out'].sum()))
# training step on the tensor-core FC kernels (gn_fc_tc.cu), two images, multi-class
setup('coco_multiclass', 2)
tr = Trainer(Gnet(80))
out = tr.step([synthetic.make_image(120, 80, image_index=1), synthetic.make_image(77, 80, image_index=2)], 1e-3)
print('tc train      ', float(out['loss_out'].sum()))
# dense IoU (symmetric kernel) and the TMA self-test
from gossipnet_b200 import ops
dd = torch.from_numpy(synthetic.make_image(333, 1)['dets']).cuda()
print('iou           ', float(ops.iou_dense(dd, dd).sum()))
mat = torch.randn(500, 64, device='cuda').to(torch.bfloat16)
wm = torch.randn(64, 64, device='cuda').to(torch.bfloat16)
dump, dout = ops.selftest_tma(mat, wm, torch.arange(128, dtype=torch.int32, device='cuda'), 0)
print('tma selftest  ', float(dout.sum()))
torch.cuda.synchronize()
print('done')
