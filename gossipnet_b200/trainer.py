"""Training step of the reference (train.py:64-77, 222-238, 316-320) on the B200
path: forward with kept activations -> DetectionMatching -> loss -> backward ->
(NCCL all-reduce of the flat gradient) -> fused Adam / Momentum update.

The reference trains one image per step (dataset.py:105).  A step here takes a
list of images; the objective is the MEAN over all images of the step (over all
ranks) of the reference's per-image loss, plus the L2 regulariser
`weight_decay * sum(W^2)/2` over the FC weights built with `weight_reg`
(everything but the predict head, network.py:261-272) applied once.  With one
image and one rank this is exactly the reference's step.

Backward follows what TF autodiff generates for network.py:
geometry features are constants (stop_gradient, :454), the pair-feature MLP
receives the sum of the gradients of all blocks, gather -> scatter-add (self
pairs excluded on the neighbour side), segment_max -> rows equal to the max
share the gradient, DetectionMatching is not differentiable (labels / weights
are constants of the step).  All arithmetic is in libgossipnet_b200.so
(csrc/gn_train.cu + the forward pieces); this file only sequences the calls.
"""
import numpy as np
import torch

from gossipnet_b200 import ops, parallel
from gossipnet_b200.nms_net.config import cfg


class Trainer(object):

    def __init__(self, net, optimizer=None, weight_decay=None, momentum=None,
                 beta1=0.9, beta2=0.999, eps=1e-8):
        self.net = net
        self.eng = net.engine
        eng = self.eng
        self.optimizer = optimizer or cfg.train.optimizer
        if self.optimizer not in ('adam', 'sgd'):
            raise ValueError('unknown optimizer {}'.format(self.optimizer))   # train.py:73
        self.momentum = cfg.train.momentum if momentum is None else momentum
        self.beta1, self.beta2, self.eps = beta1, beta2, eps
        # slim.learning.create_train_op(clip_gradient_norm=...) clips every variable's gradient
        # by its own norm (train.py:73-76); <= 0 (the shipped configs: -1) disables it
        self.clip_norm = float(cfg.train.gradient_clipping)
        wd = cfg.train.weight_decay if weight_decay is None else weight_decay
        dev = eng.device
        # flat gradient buffer + one trailing slot carrying the image count, so a
        # single all-reduce delivers both
        self.gradbuf = torch.zeros(eng.total + 1, dtype=torch.float32, device=dev)
        self.grad = self.gradbuf[:eng.total]
        self.g = dict((e.name, self.grad[e.offset:e.offset + e.size].view(*e.shape))
                      for e in eng.layout.values())
        decay = np.zeros(eng.total, dtype=np.float32)
        for e in eng.layout.values():
            if e.regularized:
                decay[e.offset:e.offset + e.size] = wd
        self.decay = torch.from_numpy(decay).to(dev)
        self.clip_table = torch.tensor([[e.offset, e.size] for e in eng.layout.values()],
                                       dtype=torch.int32, device=dev)
        self.state1 = torch.zeros(eng.total, dtype=torch.float32, device=dev)   # Adam m / momentum
        self.state2 = torch.zeros(eng.total, dtype=torch.float32, device=dev)   # Adam v
        self.global_step = 0
        self._host_images = 0      # images accumulated into gradbuf since the last update (this rank)
        self._zero_bias = torch.zeros(1024, dtype=torch.float32, device=dev)
        # every FC and its gradients on the tensor cores (gn_fc_tc.cu, bf16x3); False: the fp32
        # CUDA-core kernels (gn_fc.cu / gn_train.cu), kept as the cross-check.  Forward and
        # backward switch separately: the gradient of this piecewise-linear network (relu,
        # segment max) is discontinuous in the activations, so a 1e-5 perturbation of the
        # FORWARD flips a few selections and moves some gradients by ~1e-3 against a float64
        # oracle, while the tensor-core BACKWARD on identical activations stays within 1e-5
        # (tests/test_gpu_training.py, experiments/grad_err.py).
        self.use_tc_fwd = True
        self.use_tc_bwd = True
        self._build_fc_images()

    @property
    def use_tc(self):
        return self.use_tc_fwd and self.use_tc_bwd

    @use_tc.setter
    def use_tc(self, v):
        self.use_tc_fwd = self.use_tc_bwd = bool(v)

    def _build_fc_images(self):
        """Operand images of every FC weight for gn_fc_fwd_tc: one for y = x @ W and, where the
        input gradient is a supported shape, one for dx = dy @ W^T.  Rebuilt from the flat
        parameter buffer by ONE launch at the start of every step (the weights moved)."""
        rows, off = [], 0
        self._img_fwd, self._img_bwd = {}, {}
        for name, e in self.eng.layout.items():
            if not name.endswith('/weights') or len(e.shape) != 2:
                continue
            scope = name[:-len('/weights')]
            k, n = e.shape
            if ops.fc_tc_supported(k, n):
                kpad = (k + 15) // 16 * 16
                nbytes = 2 * (kpad // 8) * n * 16
                rows.append([e.offset, k, n, off, kpad, 0])
                self._img_fwd[scope] = (off, nbytes)
                off += nbytes
            if ops.fc_tc_supported(n, k):          # dx = dy[rows, n] @ W^T -> [rows, k]
                kpad = (n + 15) // 16 * 16
                nbytes = 2 * (kpad // 8) * k * 16
                rows.append([e.offset, k, n, off, kpad, 1])
                self._img_bwd[scope] = (off, nbytes)
                off += nbytes
        dev = self.eng.device
        self._img_table = torch.tensor(rows, dtype=torch.int32, device=dev).reshape(-1, 6)
        self._img = torch.zeros(max(off, 16), dtype=torch.uint8, device=dev)

    def _image(self, table, scope):
        off, nbytes = table[scope]
        return self._img[off:off + nbytes]

    # ---------------------------------------------------------------- helpers
    def _fc(self, x, scope, relu, residual=None, rows_dev=None):
        e = self.eng
        w = e.p[scope + '/weights']
        if self.use_tc_fwd and scope in self._img_fwd:
            return ops.fc_fwd_tc(x, self._image(self._img_fwd, scope), w.shape[0], w.shape[1],
                                 e.p[scope + '/biases'], relu, residual=residual,
                                 rows_dev=rows_dev)
        return ops.fc_fwd(x, w, e.p[scope + '/biases'], relu, residual=residual, rows_dev=rows_dev)

    def _fc_bwd(self, x, dy, scope, rows_dev=None, need_dx=True, mask=None):
        """dW, db += ; returns dx = dy @ W^T (rows beyond *rows_dev are not written).
        `mask`: the layer's activation output - its relu gradient (dy where mask > 0, else 0)
        is applied to dy inside the tensor-core kernels, or in place before the fp32 ones."""
        w = self.eng.p[scope + '/weights']
        k, n = w.shape
        tc_w = self.use_tc_bwd and k <= 256 and n in (32, 64, 128, 256)
        tc_x = self.use_tc_bwd and scope in self._img_bwd
        if mask is not None and not (tc_w and (tc_x or not need_dx)):
            ops.relu_mask(dy, mask, rows_dev=rows_dev)     # one of the consumers needs it applied
            mask = None
        if tc_w:
            ops.fc_bwd_weight_tc(x, dy, self.g[scope + '/weights'], self.g[scope + '/biases'],
                                 rows_dev=rows_dev, mask=mask)
        else:
            ops.fc_bwd_weight(x, dy, self.g[scope + '/weights'], self.g[scope + '/biases'],
                              rows_dev=rows_dev)
        if not need_dx:
            return None
        if tc_x:
            return ops.fc_fwd_tc(dy, self._image(self._img_bwd, scope), n, k, None, False,
                                 rows_dev=rows_dev, mask=mask)
        wt = ops.transpose(w)
        return ops.fc_fwd(dy, wt, self._zero_bias[:k], False, rows_dev=rows_dev)

    # -------------------------------------------------------- forward + backward
    def forward_backward(self, batches, zero_grad=True):
        """Accumulates d(sum over these images of the per-image loss)/d(theta) into
        self.grad and the image count into self.gradbuf[-1].  Returns the forward
        results (prediction, labels, weights, loss_out[B,3], ...)."""
        net, eng, g = self.net, self.eng, self.eng.g
        io = net._pack(batches, eng.device, True)
        dets, scores, classes, img_off = io['dets'], io['det_scores'], io['det_classes'], io['img_off']
        T = dets.shape[0]
        if zero_grad:
            self.gradbuf.zero_()
            self._host_images = 0
        if self.use_tc_fwd or self.use_tc_bwd:
            ops.prepare_fc_images(eng.flat, self._img_table, self._img)

        # ---- forward, keeping activations (unfused CUDA pieces) ---------------------
        max_img = int(np.max(np.diff(io['img_off_host'])))
        row_ptr, num_pairs, pair_c, pair_n, pair_iou, cap = eng.neighbors(dets, img_off, max_img)
        P = int(num_pairs.item())
        if P > cap:   # grow and redo the fill
            eng.capacity = int(P * 1.25) + 256
            row_ptr, num_pairs, pair_c, pair_n, pair_iou, cap = eng.neighbors(dets, img_off,
                                                                              max_img)
        cls = classes if eng.multiclass else None
        raw = ops.pair_geometry(dets, scores, cls, pair_c, pair_n, pair_iou, num_pairs, cap,
                                eng.num_classes, g['pw_feat_multiplyer'])
        pw_acts = [raw]
        for i in range(1, g['num_pwfeat_fc'] + 1):
            pw_acts.append(self._fc(pw_acts[-1], 'gnet/pw_feats/fc%d' % i, True, rows_dev=num_pairs))
        pw = pw_acts[-1]
        w, r, f = pw.shape[1], g['reduced_dim'], g['pairfeat_dim']

        im_acts, im_scopes = None, []
        if eng.imfeats:
            # image-feature head (network.py:223-240); the feature maps are inputs of the
            # step (the ResNet that makes them is not part of this package), so the gradient
            # stops at the ROI-pooled features
            start, roifeats, _, _ = eng.image_features(
                dets, io['img_off_host'], net._pack_imfeats(batches),
                torch.empty((T, g['shortcut_dim']), dtype=torch.float32, device=eng.device))
            x0 = roifeats.reshape(T, -1)
            scope = 'gnet/reduce_imfeats/fully_connected'
            im_acts = [x0]
            if eng.imfeat_dim > 0:
                im_acts.append(self._fc(x0, scope, True))
                im_scopes.append(scope)
                scope += '_1'
            im_scopes.append(scope)
            feats = self._fc(im_acts[-1], scope, True)
            im_acts.append(feats)
        else:
            feats = torch.zeros((T, g['shortcut_dim']), dtype=torch.float32, device=eng.device)
        tape = []
        for b in range(1, g['num_blocks'] + 1):
            s = 'gnet/block%d/' % b
            red = self._fc(feats, s + 'reduce_dim', True)
            nred = self._fc(feats, s + 'reduce_dim_neighbor', True) if g['neighbor_feats'] else red
            x = ops.block_gather_concat(pw, red, nred, pair_c, pair_n, num_pairs, cap)
            hs = []
            h = x
            for i in range(1, g['num_block_pw_fc'] + 1):
                h = self._fc(h, s + 'pw_fc%d' % i, True, rows_dev=num_pairs)
                hs.append(h)
            # x stays alive for the backward (pw_fc1's weight gradient needs it): 119 MB per
            # block at P = 311 k is nothing on 180 GB, and re-gathering cost 79 us per block
            pooled = ops.segment_max(h, row_ptr, T)
            ds = [pooled]
            for i in range(1, g['num_block_fc']):
                ds.append(self._fc(ds[-1], s + 'fc%d' % i, True))
            out = self._fc(ds[-1], s + 'fc%d' % g['num_block_fc'], True, residual=feats)
            tape.append((feats, red, nred, hs, ds, out, x))
            feats = out
        pred_acts = [feats]
        for i in range(1, g['num_predict_fc']):
            pred_acts.append(self._fc(pred_acts[-1], 'gnet/predict/fc%d/fully_connected' % i, False))
        prediction = self._fc(pred_acts[-1], 'gnet/predict/logits/fully_connected', False).view(-1)

        res = eng.matching_and_loss(prediction, dets, classes, img_off, io['img_off_host'],
                                    io['gt_boxes'], io['gt_crowd'], io['gt_classes'],
                                    io['gt_off_host'], net.class_weights, want_grad=True)
        res.update(prediction=prediction, P=P, num_images=len(batches))

        # ---- backward --------------------------------------------------------------
        d = res['dlogit'].view(-1, 1).contiguous()
        d = self._fc_bwd(pred_acts[-1], d, 'gnet/predict/logits/fully_connected')
        for i in range(g['num_predict_fc'] - 1, 0, -1):
            d = self._fc_bwd(pred_acts[i - 1], d, 'gnet/predict/fc%d/fully_connected' % i)
        dfeats = d
        dpw = torch.zeros((cap, w), dtype=torch.float32, device=eng.device)
        for b in range(g['num_blocks'], 0, -1):
            s = 'gnet/block%d/' % b
            feats_in, red, nred, hs, ds, out, x = tape[b - 1]
            tape[b - 1] = None                               # free this block's activations
            ops.relu_mask(dfeats, out)                      # shortcut relu (network.py:407-408)
            dd = self._fc_bwd(ds[-1], dfeats, s + 'fc%d' % g['num_block_fc'])
            for i in range(g['num_block_fc'] - 1, 0, -1):
                ops.relu_mask(dd, ds[i])
                dd = self._fc_bwd(ds[i - 1], dd, s + 'fc%d' % i)
            dh = torch.empty_like(hs[-1]) if hs else None
            if hs:
                ops.segment_max_bwd(hs[-1], ds[0], dd, row_ptr, dh)
                for i in range(g['num_block_pw_fc'], 0, -1):
                    dh = self._fc_bwd(hs[i - 2] if i > 1 else x, dh, s + 'pw_fc%d' % i,
                                      rows_dev=num_pairs, mask=hs[i - 1])
            else:
                dh = torch.empty_like(x)
                ops.segment_max_bwd(x, ds[0], dd, row_ptr, dh)
            dred = torch.zeros_like(red)
            dnred = torch.zeros_like(red) if g['neighbor_feats'] else dred
            ops.gather_concat_bwd(dh, w, r, pair_c, pair_n, row_ptr, T, num_pairs, cap, dpw,
                                  dred, dnred)
            ops.relu_mask(dred, red)
            first_needs_dx = b > 1 or im_acts is not None
            dx = self._fc_bwd(feats_in, dred, s + 'reduce_dim', need_dx=first_needs_dx)
            if g['neighbor_feats']:
                ops.relu_mask(dnred, nred)
                dxn = self._fc_bwd(feats_in, dnred, s + 'reduce_dim_neighbor',
                                   need_dx=first_needs_dx)
                if dxn is not None:
                    ops.add_inplace(dfeats, dxn)
            # without image features block 1's input is the constant zero start feature
            # (network.py:241-246) and nothing flows further
            if dx is not None:
                ops.add_inplace(dfeats, dx)
        if im_acts is not None:       # dfeats = d loss / d start features
            d = dfeats
            for i in range(len(im_scopes), 0, -1):
                ops.relu_mask(d, im_acts[i])
                d = self._fc_bwd(im_acts[i - 1], d, im_scopes[i - 1], need_dx=i > 1)
        # pair-feature MLP: sum of the gradients of all blocks; raw features are constants
        d = dpw
        for i in range(g['num_pwfeat_fc'], 0, -1):
            d = self._fc_bwd(pw_acts[i - 1], d, 'gnet/pw_feats/fc%d' % i, rows_dev=num_pairs,
                             need_dx=i > 1, mask=pw_acts[i])
        self.gradbuf[-1] += float(len(batches))
        self._host_images += len(batches)
        return res

    # ------------------------------------------------------------------- update
    def apply_gradients(self, lr):
        """All-reduce (sum) the flat gradient + image count, then the optimizer update
        with grad_scale = 1 / (images of the step over all ranks)."""
        parallel.allreduce_sum_(self.gradbuf)
        if parallel.world() > 1:
            n_images = float(self.gradbuf[-1].item())   # the other ranks' shards: read back (one sync)
        else:
            n_images = float(self._host_images)         # single process: known here, no host sync
        self._host_images = 0
        scale = 1.0 / max(n_images, 1.0)
        self.global_step += 1
        self.eng.weights_version += 1        # the optimizer kernel writes the flat buffer in place
        decay = self.decay
        if self.clip_norm > 0:
            # the clipped gradient of the total loss (data term + l2 regularizer) replaces grad
            ops.clip_gradients(self.grad, self.eng.flat, self.decay, self.clip_table, scale,
                               self.clip_norm)
            scale, decay = 1.0, None
        if self.optimizer == 'adam':
            ops.adam_step(self.eng.flat, self.grad, self.state1, self.state2, decay, lr,
                          self.beta1, self.beta2, self.eps, self.global_step, scale)
        else:
            ops.momentum_step(self.eng.flat, self.grad, self.state1, decay, lr,
                              self.momentum, scale)
        return n_images

    def step(self, batches, lr):
        """One optimizer step on this rank's images (possibly none: a rank whose
        shard of a small step is empty still joins the all-reduce)."""
        if len(batches) == 0:
            self.gradbuf.zero_()
            self._host_images = 0
            res = {'num_images': 0, 'loss_out': None}
        else:
            res = self.forward_backward(batches)
        res['images_in_step'] = self.apply_gradients(lr)
        return res

    def regularization_loss(self):
        """weight_decay * sum(W^2) / 2 over the regularised weights: the term
        tf.contrib.losses.get_total_loss() adds to the data loss (train.py:237).
        Display only; the update applies its gradient inside the optimizer kernel."""
        flat = self.eng.flat
        return float(0.5 * torch.sum(self.decay * flat * flat).item())

    # ------------------------------------------------------- checkpoint / resume
    def state_dict(self):
        return {'params': self.eng.flat.detach().cpu(), 'state1': self.state1.cpu(),
                'state2': self.state2.cpu(), 'global_step': self.global_step,
                'optimizer': self.optimizer}

    def load_state_dict(self, sd):
        self.eng.flat.copy_(sd['params'].to(self.eng.device))
        self.state1.copy_(sd['state1'].to(self.eng.device))
        self.state2.copy_(sd['state2'].to(self.eng.device))
        self.global_step = int(sd['global_step'])


class LearningRate(object):
    """cfg.train.lr_multi_step schedule, same (stateful) behaviour as train.py:26-37."""

    def __init__(self):
        self.steps = cfg.train.lr_multi_step
        self.current_step = 0

    def get_lr(self, iter):
        if self.current_step >= len(self.steps):
            return self.steps[-1][1]
        lr = self.steps[self.current_step][1]
        if iter == self.steps[self.current_step][0]:
            self.current_step += 1
        return lr
