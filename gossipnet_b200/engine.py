"""Batched Gnet forward on one B200: the host-side sequencing of the C-ABI calls.

One `GnetEngine` owns the flat parameter buffer and a grow-only workspace and
runs the whole hot path for a BATCH of images whose
detections are concatenated (`img_off[num_images+1]`): neighbor build ->
pair-feature MLP -> `num_blocks` blocks -> predict head (-> det/GT IoU ->
DetectionMatching -> loss when ground truth is given).  Every step is one call
into libgossipnet_b200.so on torch's current stream; nothing here computes.

The pair count P is data dependent.  It stays on the device (`row_ptr[T]`) and
every pair kernel reads it there, so a forward never synchronises: buffers are
sized by a `capacity` the engine learns on its first call (one sync, then
grow-only with a 25 % margin) and an overflow flag is checked by the caller
together with the result it reads back anyway (`check_overflow`).  That makes
`forward` CUDA-graph capturable (bench.py does so).

Reference: nms_net/network.py:148-322 (graph construction order), :344-409
(block), :257-273 (predict head), :275-314 (loss).
"""
import numpy as np
import torch

from gossipnet_b200 import ops
from gossipnet_b200 import params as P


class CapacityOverflow(RuntimeError):
    """More neighbor pairs than the workspace was sized for; the engine has
    grown its capacity, call forward again."""


class GnetEngine(object):

    def __init__(self, num_classes, cfg, device='cuda', flat_params=None, seed=None):
        if not torch.cuda.is_available():
            raise RuntimeError('gossipnet_b200 needs a CUDA device: the hot path has no CPU '
                               'fallback')
        ops._lib.load()  # fail loudly now if the extension is not built
        self.device = torch.device(device)
        self.num_classes = int(num_classes)
        self.multiclass = self.num_classes > 1
        g = cfg.gnet
        # snapshot of the hyper-parameters the path reads (cfg is a mutable global)
        self.g = dict((k, g[k]) for k in (
            'neighbor_thresh', 'shortcut_dim', 'num_blocks', 'reduced_dim', 'pairfeat_dim',
            'num_block_pw_fc', 'num_block_fc', 'num_predict_fc', 'predict_fc_dim',
            'neighbor_feats', 'num_pwfeat_fc', 'pwfeat_dim', 'pwfeat_narrow_dim',
            'pw_feat_multiplyer'))
        # image-feature head (network.py:223-240): ROI-pooled crops of a feature map
        if g.compute_dtype not in ('fp32', 'bf16'):
            raise ValueError('cfg.gnet.compute_dtype must be fp32 or bf16, got %r' % (g.compute_dtype,))
        self.bf16 = g.compute_dtype == 'bf16'
        self.imfeats = bool(g.imfeats)
        self.imfeat_dim = int(g.imfeat_dim)
        self.crop = (int(cfg.imfeat_crop_height), int(cfg.imfeat_crop_width))
        self.imfeat_stride = 16          # get_resnet: output_stride=16 (network.py:57,63)
        self.normalize_loss = bool(cfg.train.normalize_loss)
        self.loss_multiplyer = float(cfg.train.loss_multiplyer)
        self.layout, self.total = P.param_layout(num_classes, cfg)
        if flat_params is None:
            flat_params = P.init_flat(self.layout, self.total, cfg, seed=seed)
        if isinstance(flat_params, np.ndarray):
            flat_params = torch.from_numpy(np.ascontiguousarray(flat_params, dtype=np.float32))
        if flat_params.numel() != self.total:
            raise ValueError('flat parameter buffer has %d elements, layout needs %d'
                             % (flat_params.numel(), self.total))
        self.flat = flat_params.to(self.device, torch.float32).contiguous()
        self.p = P.views(self.layout, self.flat)

        self.raw_width = P.raw_pairfeat_width(num_classes)
        self.pw_width = g.pwfeat_narrow_dim if g.num_pwfeat_fc > 0 else self.raw_width
        # the fused kernels are built for the shipped shapes; anything else runs
        # through the unfused CUDA pieces (same results, more HBM traffic)
        self.fused_pw = (g.num_pwfeat_fc == 3 and g.pwfeat_dim == 256
                         and g.pwfeat_narrow_dim == 32)
        self.fused_block = (self.pw_width == 32 and g.reduced_dim == 32
                            and g.pairfeat_dim == 64 and g.num_block_pw_fc == 2)
        # detection-level layers fused across the block boundary (gn_block_det_fwd)
        self.fused_det = (self.fused_block and g.shortcut_dim == 128 and g.num_block_fc == 2
                          and not g.neighbor_feats)
        self.capacity = 0
        self._ws = {}
        self.keep_block_feats = False
        self.use_fused = True
        self.use_tensor_cores = True   # False: fp32 FFMA variants of the fused kernels
        # 'hl': the reference's K = 96 formulation on bf16 hi/lo feature rows
        # (gn_block_tc.cu); 'ab': first pair FC split into per-detection halves
        # (gn_block_ab.cu).  Measured on the bench workload: 256 vs 323 us per block for
        # the pair kernel (+14 us in the det kernel): 'ab' halves the shared-memory
        # traffic but doubles the random L2 gather bytes per pair (2 x 256 B fp32 rows
        # instead of 2 x 128 B), whose latency lands in the epilogue.  Kept selectable.
        # 'pipe': the 'hl' arithmetic as a warp-specialised pipeline (gn_block_pipe.cu).
        # 'tma': operands fed by tensor-map TMA (tile load of the pair's own bf16 hi|lo pw row,
        # gather4 of the neighbor's reduced row), the detection-level third of pw_fc1 hoisted
        # into the det kernel (gn_block_tma.cu).
        self.pair_mode = 'tma'
        # detection-level kernel of the 'tma' path: tile transfers on the copy engine
        # (gn_det_tma.cu); False: the eight-warp kernel of gn_det_tc.cu (same results)
        self.det_tma = True
        # fp32 pw_feats next to the bf16 (hi|lo) operand rows the 'tma' pair stage consumes
        # (the `pw_feats` attribute of the reference surface); False on the inference hot path
        self.want_pw_f32 = True
        # bumped whenever a workspace buffer is reallocated: captured CUDA graphs hold raw
        # pointers into the workspace and must be dropped when this changes (session.py)
        self.ws_generation = 0
        # Everything derived from the weights alone (operand images, the folded predict head)
        # is rebuilt only when the weights changed: `flat._version` counts torch-side in-place
        # writes (load_state_dict, Variable.value.copy_), `weights_version` is bumped by callers
        # that write through the C ABI (the optimizer step).  A captured CUDA graph does not
        # contain these launches; InferenceSession refreshes them before a replay when the key
        # moved (refresh_weight_images).
        self.weights_version = 0
        self._derived_key = {}
        # the predict head's hidden layers are linear (network.py:263): apply it as one folded
        # affine map instead of three FC launches (False: the staged FCs)
        self.collapse_predict = True
        # neighbor build through per-row hit masks (division-free threshold test in the count
        # pass, exact IoU only for the hits in the fill pass); False: both passes recompute
        self.use_neighbor_masks = True

    # ------------------------------------------------------- weight-derived buffers
    def weights_key(self):
        return (self.flat._version, self.weights_version, self.pair_mode, self.bf16)

    def _stale(self, what):
        """True (and marks `what` fresh) when the weights changed since `what` was last built."""
        key = self.weights_key()
        if self._derived_key.get(what) == key:
            return False
        self._derived_key[what] = key
        return True

    def refresh_weight_images(self):
        """Rebuild every weight-derived buffer of the fused forward now (on the current stream)
        if the weights moved; returns True when anything was launched.  For callers that
        replay a captured forward."""
        done = False
        if self.fused_det and self.use_fused and self.use_tensor_cores and self.g['num_blocks'] > 0:
            done |= self._prepare_block_images()
        if self.use_fused and self.collapse_predict:
            done |= self._prepare_predict()
        return done

    # ------------------------------------------------------------------ workspace
    def _buf(self, name, shape, dtype=torch.float32):
        n = int(np.prod(shape))
        t = self._ws.get(name)
        if t is None or t.numel() < n or t.dtype != dtype:
            t = torch.empty(max(n, 1), dtype=dtype, device=self.device)
            self._ws[name] = t
            self.ws_generation += 1
        return t[:n].view(*shape)

    def _ensure_capacity(self, num_pairs_dev, num_dets):
        if self.capacity == 0:
            p = int(num_pairs_dev.item())  # the one sync of the engine's lifetime
            self.capacity = max(1024, int(p * 1.25) + 256)
        return self.capacity

    def check_overflow(self):
        """Call after reading a result back: raises CapacityOverflow (after
        growing) if the last forward saw more pairs than it had room for."""
        p = int(self._last_num_pairs.item())
        if p > self._last_capacity:
            self.capacity = int(p * 1.25) + 256
            raise CapacityOverflow('P=%d > capacity=%d' % (p, self._last_capacity))
        return p

    # -------------------------------------------------------------------- forward
    def neighbors(self, dets, img_off, max_img=None):
        """A3. -> (row_ptr[T+1], num_pairs_dev[1], pair_c, pair_n, pair_iou, capacity).
        `max_img` = detections of the largest image (host int): sizes the per-row hit masks
        of the mask variant; None reads it back from img_off (one sync)."""
        T = dets.shape[0]
        degree = self._buf('degree', (T,), torch.int32)
        row_ptr = self._buf('row_ptr', (T + 1,), torch.int32)
        masks = None
        if self.use_neighbor_masks:
            if max_img is None:
                max_img = int((img_off[1:] - img_off[:-1]).max().item())
            stride = max(1, (int(max_img) + 31) // 32)
            # the dense-proposal stress image (N = 10 000) needs 313 words x 10 000 rows = 12.5 MB
            masks = self._buf('nb_masks', (T, stride), torch.int32)
            ops.neighbor_count_masks(dets, img_off, self.g['neighbor_thresh'], stride, degree, masks)
        else:
            ops.neighbor_count(dets, img_off, self.g['neighbor_thresh'], degree)
        ops.exclusive_scan(degree, row_ptr)
        num_pairs = row_ptr[T:T + 1]
        cap = self._ensure_capacity(num_pairs, T)
        if not (self.fused_block and self.use_fused):
            # the unfused pieces index pair rows through row_ptr, so they need
            # P <= capacity BEFORE going on (not the hot path: one sync is fine)
            p = int(num_pairs.item())
            if p > cap:
                cap = self.capacity = int(p * 1.25) + 256
        pair_c = self._buf('pair_c', (cap,), torch.int32)
        pair_n = self._buf('pair_n', (cap,), torch.int32)
        pair_iou = self._buf('pair_iou', (cap,), torch.float32)
        if masks is not None:
            ops.neighbor_fill_masks(dets, img_off, row_ptr, cap, masks, masks.shape[1], pair_c,
                                    pair_n, pair_iou)
        else:
            ops.neighbor_fill(dets, img_off, self.g['neighbor_thresh'], row_ptr, cap, pair_c,
                              pair_n, pair_iou, None)
        self._last_num_pairs, self._last_capacity = num_pairs, cap
        return row_ptr, num_pairs, pair_c, pair_n, pair_iou, cap

    def _fc(self, x, scope, relu, residual=None, out=None, rows_dev=None):
        return ops.fc_fwd(x, self.p[scope + '/weights'], self.p[scope + '/biases'], relu,
                          residual=residual, out=out, rows_dev=rows_dev)

    def pair_features(self, dets, scores, classes, pair_c, pair_n, pair_iou, num_pairs, cap):
        """A4 + A5 -> pw_feats[cap, pw_width] (rows >= P undefined)."""
        g = self.g
        cls = classes if self.multiclass else None
        mult = g['pw_feat_multiplyer']
        if g['num_pwfeat_fc'] > 0 and self.fused_pw and self.use_fused:
            s = 'gnet/pw_feats/fc%d/'
            w = [self.p[(s % i) + k] for i in (1, 2, 3) for k in ('weights', 'biases')]
            if 'wprep' not in self._ws:
                self._ws['wprep'] = torch.empty(int(ops._lib.load().gn_pwfeat_prep_bytes()),
                                                dtype=torch.uint8, device=self.device)
                self.ws_generation += 1
            if self._tma_path():
                # bf16 (hi | lo) operand rows for the TMA-fed pair stage; fp32 rows on request
                self._pw_hl = self._buf('pw_hl', (cap, 64), torch.bfloat16)
                out = self._buf('pw', (cap, 32)) if self.want_pw_f32 else None
                return ops.pwfeat_mlp_fwd(dets, scores, cls, pair_c, pair_n, pair_iou, num_pairs,
                                          cap, self.num_classes, mult, *w, out=out,
                                          wprep=self._ws['wprep'], bf16=self.bf16,
                                          out_hl=self._pw_hl, want_f32=self.want_pw_f32)
            out = self._buf('pw', (cap, 32))
            return ops.pwfeat_mlp_fwd(dets, scores, cls, pair_c, pair_n, pair_iou, num_pairs, cap,
                                      self.num_classes, mult, *w, out=out,
                                      ffma=not self.use_tensor_cores, wprep=self._ws['wprep'],
                                      bf16=self.bf16)
        raw = self._buf('pw_raw', (cap, self.raw_width))
        ops.pair_geometry(dets, scores, cls, pair_c, pair_n, pair_iou, num_pairs, cap,
                          self.num_classes, mult, raw)
        x = raw
        for i in range(1, g['num_pwfeat_fc'] + 1):
            width = g['pwfeat_dim'] if i < g['num_pwfeat_fc'] else g['pwfeat_narrow_dim']
            y = self._buf('pw_fc%d' % (i % 2), (cap, width))
            x = self._fc(x, 'gnet/pw_feats/fc%d' % i, True, out=y, rows_dev=num_pairs)
        return x

    def _tma_path(self):
        """True when the forward runs the TMA-fed pair stage (needs every fused piece)."""
        return (self.pair_mode == 'tma' and self.fused_pw and self.fused_det and self.use_fused
                and self.use_tensor_cores and self.g['num_blocks'] > 0
                and self.g['num_pwfeat_fc'] > 0)

    def block(self, b, infeats, row_ptr, pair_c, pair_n, num_pairs, cap, pw, out):
        """A7: nms_net/network.py:344-409."""
        g = self.g
        T = infeats.shape[0]
        s = 'gnet/block%d/' % b
        red = self._fc(infeats, s + 'reduce_dim', True, out=self._buf('red', (T, g['reduced_dim'])))
        nred = red
        if g['neighbor_feats']:
            nred = self._fc(infeats, s + 'reduce_dim_neighbor', True,
                            out=self._buf('nred', (T, g['reduced_dim'])))
        f = g['pairfeat_dim']
        pooled = self._buf('pooled', (T, f))
        if self.fused_block and self.use_fused:
            pooled.zero_()
            ops.block_pair_fwd(pw, red, nred, pair_c, pair_n, num_pairs, cap,
                               self.p[s + 'pw_fc1/weights'], self.p[s + 'pw_fc1/biases'],
                               self.p[s + 'pw_fc2/weights'], self.p[s + 'pw_fc2/biases'], pooled,
                               ffma=not self.use_tensor_cores)
        else:
            x = ops.block_gather_concat(pw, red, nred, pair_c, pair_n, num_pairs, cap,
                                        self._buf('pairx', (cap, pw.shape[1] + 2 * g['reduced_dim'])))
            for i in range(1, g['num_block_pw_fc'] + 1):
                x = self._fc(x, s + 'pw_fc%d' % i, True, out=self._buf('pairh%d' % (i % 2), (cap, f)),
                             rows_dev=num_pairs)
            if g['num_block_pw_fc'] == 0:
                f = x.shape[1]
                pooled = self._buf('pooled', (T, f))
            ops.segment_max(x, row_ptr, T, pooled)
        x = pooled
        for i in range(1, g['num_block_fc']):
            x = self._fc(x, s + 'fc%d' % i, True, out=self._buf('deth%d' % (i % 2), (T, g['pairfeat_dim'])))
        # last FC has no activation; shortcut: relu(infeats + feats) (:399-408)
        return self._fc(x, s + 'fc%d' % g['num_block_fc'], True, residual=infeats, out=out)

    def _operand_images(self):
        """Per-block operand images (bf16 hi/lo K-major tiles of the block weights) and the
        table that lets ONE gn_prepare_operands launch rebuild them from the flat buffer.
        pair_mode 'ab': pair image = [pw_fc1[0:32]^T, pw_fc2^T]; det image b = [fc1, fc2 of
        block b, reduce_dim and (pw_fc1[32:64] | pw_fc1[64:96])^T of block b+1].
        pair_mode 'hl': pair image = [pw_fc1^T (K = 96), pw_fc2^T]; det image without the
        last part."""
        ab = self.pair_mode in ('ab', 'tma')
        key = 'wimg_' + ('ab' if ab else 'hl')
        if key in self._ws:
            return self._ws[key]
        lib = ops._lib.load()
        pair_b = int(lib.gn_block_pair_ab_image_bytes() if ab else lib.gn_block_pair_image_bytes())
        det_b = int(lib.gn_block_det_image_bytes())
        nb = self.g['num_blocks']
        rows, off = [], 0
        pair_off, det_off = [], []

        def add(name, dst, row0=0, k=None, pitch=0, col_off=0):
            """rows [row0, row0+k) of weight `name` -> hi tile at dst, lo tile right after
            (tile bytes = (k/8) * max(pitch, n*16)); returns the end offset."""
            e = self.layout[name + '/weights']
            kk, n = e.shape
            k = kk - row0 if k is None else k
            tile = (k // 8) * (pitch if pitch else n * 16)
            rows.append([e.offset + row0 * n, k, n, dst + col_off, dst + tile + col_off, pitch])
            return dst + 2 * tile

        for b in range(1, nb + 1):
            s = 'gnet/block%d/' % b
            pair_off.append(off)
            o = add(s + 'pw_fc1', off, 0, 32 if ab else None)
            o = add(s + 'pw_fc2', o)
            assert o - off == pair_b, (o - off, pair_b)
            off = o
        for b in range(0, nb + 1):     # det image b: fc1/fc2 of block b, reduce_dim (+AB) of b+1
            det_off.append(off)
            o = off
            if b >= 1:
                o = add('gnet/block%d/fc1' % b, o)
                o = add('gnet/block%d/fc2' % b, o)
            else:
                o += (8 * 64 + 8 * 128) * 16 * 2
            if b + 1 <= nb:
                o = add('gnet/block%d/reduce_dim' % (b + 1), o)
                if ab:   # two [32, 64] row blocks side by side in one N = 128 tile
                    name = 'gnet/block%d/pw_fc1' % (b + 1)
                    add(name, o, 32, 32, pitch=128 * 16, col_off=0)
                    add(name, o, 64, 32, pitch=128 * 16, col_off=64 * 16)
            off += det_b
        table = torch.tensor(rows, dtype=torch.int32, device=self.device)
        image = torch.zeros(off, dtype=torch.uint8, device=self.device)
        self._ws[key] = (image, table, (pair_off, det_off, pair_b, det_b))
        self.ws_generation += 1
        self._derived_key.pop('block_images', None)      # a fresh buffer: rebuild
        return self._ws[key]

    def _tma_images(self):
        """Swizzled operand images of the TMA-fed pair stage (one per block) + the offset
        table gn_prepare_pair_tma_image rebuilds them from."""
        if 'wimg_tma' not in self._ws:
            nb = self.g['num_blocks']
            rows = [[self.layout['gnet/block%d/pw_fc1/weights' % b].offset,
                     self.layout['gnet/block%d/pw_fc2/weights' % b].offset]
                    for b in range(1, nb + 1)]
            table = torch.tensor(rows, dtype=torch.int32, device=self.device)
            image = torch.zeros(nb * ops.pair_tma_image_bytes(), dtype=torch.uint8,
                                device=self.device)
            self._ws['wimg_tma'] = (image, table)
            self.ws_generation += 1
            self._derived_key.pop('block_images', None)
        return self._ws['wimg_tma']

    def _prepare_block_images(self):
        """Operand images of all blocks from the flat parameter buffer (two launches), skipped
        while the weights have not changed."""
        if not self._stale('block_images'):
            return False
        image, table, _ = self._operand_images()
        ops.prepare_operands(self.flat, table, image)
        if self._tma_path():
            tma_image, tma_table = self._tma_images()
            ops.prepare_pair_tma_image(self.flat, tma_table, tma_image)
        return True

    def _blocks_fused(self, feats, pair_c, pair_n, num_pairs, cap, pw, block_feats):
        """All blocks with two launches each: the tensor-core pair stage and the fused
        detection-level kernel (fc1, fc2, shortcut of block b + reduce_dim of block b+1
        + the per-detection halves of block b+1's first pair FC); weights travel as
        operand images prepared by one launch per forward."""
        g, p = self.g, self.p
        T, d = feats.shape
        if self.bf16 and self.pair_mode not in ('pipe', 'tma'):
            raise ValueError("cfg.gnet.compute_dtype = 'bf16' needs pair_mode 'tma' or 'pipe'")
        ab_mode = self.pair_mode == 'ab'
        tma_mode = self._tma_path()
        if self.pair_mode == 'tma' and not tma_mode:
            raise ValueError("pair_mode 'tma' needs the fused pair-feature MLP (shipped shapes)")
        pooled = self._buf('pooled', (T, g['pairfeat_dim']))
        pooled.zero_()   # every det launch re-zeroes it; this covers a dirty workspace
        if ab_mode:
            inter = self._buf('ab', (T, 2 * g['pairfeat_dim']))
        if tma_mode:
            inter = self._buf('u', (T, g['pairfeat_dim']))     # U = red @ pw_fc1[32:64] + b
        if tma_mode:
            # reduced features as bf16 (hi | lo) rows + the all-zero row a self pair gathers
            red_all = self._buf('red_hl', (T + 1, 2 * g['reduced_dim']), torch.bfloat16)
            red_all[T].zero_()
            inter_hl = red_all[:T]
            tma_image, tma_table = self._tma_images()
            tma_b = ops.pair_tma_image_bytes()
        elif not ab_mode:
            inter = self._buf('red_hl', (T, 2 * g['reduced_dim']), torch.bfloat16)
        image, table, (pair_off, det_off, pair_b, det_b) = self._operand_images()
        self._prepare_block_images()
        nb = g['num_blocks']

        def det(b, pooled_in, feats_in, out):
            """det kernel after block b (b = 0: only the reduce of block 1)."""
            s = 'gnet/block%d/' % b
            nxt = 'gnet/block%d/' % (b + 1)
            last = b == nb
            if tma_mode and self.det_tma:
                ops.block_det_fwd_tma(
                    pooled_in, feats_in, image[det_off[b]:det_off[b] + det_b],
                    p[s + 'fc1/biases'] if b >= 1 else None, p[s + 'fc2/biases'] if b >= 1 else None,
                    None if last else p[nxt + 'reduce_dim/biases'], out,
                    None if last else inter_hl, None if last else p[nxt + 'pw_fc1/biases'],
                    None if last else inter, bf16=self.bf16)
                return
            if tma_mode:
                ops.block_det_fwd_img_u(
                    pooled_in, feats_in, image[det_off[b]:det_off[b] + det_b],
                    p[s + 'fc1/biases'] if b >= 1 else None, p[s + 'fc2/biases'] if b >= 1 else None,
                    None if last else p[nxt + 'reduce_dim/biases'], out,
                    None if last else inter_hl, None if last else p[nxt + 'pw_fc1/biases'],
                    None if last else inter, bf16=self.bf16)
                return
            ops.block_det_fwd_img(
                pooled_in, feats_in, image[det_off[b]:det_off[b] + det_b],
                p[s + 'fc1/biases'] if b >= 1 else None, p[s + 'fc2/biases'] if b >= 1 else None,
                None if last else p[nxt + 'reduce_dim/biases'], feats_out=out,
                red_hl=None if (last or ab_mode) else inter,
                b_ab=p[nxt + 'pw_fc1/biases'] if (ab_mode and not last) else None,
                ab_out=inter if (ab_mode and not last) else None, bf16=self.bf16)

        det(0, None, feats, None)
        for b in range(1, nb + 1):
            s = 'gnet/block%d/' % b
            wimg = image[pair_off[b - 1]:pair_off[b - 1] + pair_b]
            if tma_mode:
                ops.block_pair_fwd_tma(self._pw_hl, red_all, T, inter, pair_c, pair_n, num_pairs,
                                       cap, p[s + 'pw_fc2/biases'],
                                       tma_image[(b - 1) * tma_b:b * tma_b], pooled,
                                       bf16=self.bf16)
            elif ab_mode:
                ops.block_pair_fwd_ab(pw, inter, pair_c, pair_n, num_pairs, cap,
                                      p[s + 'pw_fc2/biases'], wimg, pooled)
            elif self.pair_mode == 'pipe':
                ops.block_pair_fwd_pipe(pw, inter, pair_c, pair_n, num_pairs, cap,
                                        p[s + 'pw_fc1/biases'], p[s + 'pw_fc2/biases'], wimg, pooled,
                                        bf16=self.bf16)
            else:
                ops.block_pair_fwd(pw, inter, inter, pair_c, pair_n, num_pairs, cap,
                                   None, p[s + 'pw_fc1/biases'], None, p[s + 'pw_fc2/biases'],
                                   pooled, wimg=wimg)
            out = self._buf('feats%d' % (b % 2), (T, d))
            det(b, pooled, feats, out)
            feats = out
            if block_feats is not None:
                block_feats.append(feats.clone())
        return feats

    def predict(self, feats):
        """A8: two LINEAR layers then the logit (network.py:257-273)."""
        g = self.g
        T = feats.shape[0]
        if self.use_fused and self.collapse_predict:
            return self._predict_collapsed(feats)
        x = feats
        for i in range(1, g['num_predict_fc']):
            x = self._fc(x, 'gnet/predict/fc%d/fully_connected' % i, False,
                         out=self._buf('pred%d' % (i % 2), (T, g['predict_fc_dim'])))
        out = self._fc(x, 'gnet/predict/logits/fully_connected', False,
                       out=self._buf('logits', (T, 1)))
        return out.view(-1)

    def _prepare_predict(self):
        """The head's hidden layers are linear, so it is ONE affine map: fold the chain into
        w_eff / b_eff (one small launch), skipped while the weights have not changed."""
        g = self.g
        if 'pred_table' not in self._ws:
            names = ['gnet/predict/fc%d/fully_connected' % i for i in range(1, g['num_predict_fc'])]
            names.append('gnet/predict/logits/fully_connected')
            rows = []
            for nm in names:
                w, b = self.layout[nm + '/weights'], self.layout[nm + '/biases']
                rows.append([w.offset, b.offset, w.shape[0], w.shape[1]])
            self._ws['pred_table'] = torch.tensor(rows, dtype=torch.int32, device=self.device)
            self._ws['pred_maxdim'] = max(max(r[2], r[3]) for r in rows)
            self._ws['pred_weff'] = torch.empty(g['shortcut_dim'], dtype=torch.float32, device=self.device)
            self._ws['pred_beff'] = torch.empty(1, dtype=torch.float32, device=self.device)
            self._derived_key.pop('predict', None)
        if not self._stale('predict'):
            return False
        table, md = self._ws['pred_table'], self._ws['pred_maxdim']
        ops.predict_collapse(self.flat, table, md, self._buf('pred_scratch', (2 * md,)),
                             self._ws['pred_weff'], self._ws['pred_beff'])
        return True

    def _predict_collapsed(self, feats):
        """One dot product per row with the folded head (network.py:257-273)."""
        self._prepare_predict()
        return ops.rowdot_fwd(feats, self._ws['pred_weff'], self._ws['pred_beff'],
                              self._buf('logits', (feats.shape[0],)))

    def _empty_result(self):
        """A batch without a single detection (every image empty): nothing to launch."""
        dev, f32, i32 = self.device, torch.float32, torch.int32
        zero = torch.zeros(1, dtype=i32, device=dev)
        self._last_num_pairs, self._last_capacity = zero, max(self.capacity, 1)
        e = lambda shape, dt=f32: torch.empty(shape, dtype=dt, device=dev)
        res = dict(prediction=e((0,)), row_ptr=zero, num_pairs=zero, pair_c=e((0,), i32),
                   pair_n=e((0,), i32), pair_iou=e((0,)), pw_feats=e((0, self.pw_width)),
                   feats=e((0, self.g['shortcut_dim'])), capacity=0)
        if self.keep_block_feats:
            res['block_feats'] = [e((0, self.g['shortcut_dim']))
                                  for _ in range(self.g['num_blocks'] + 1)]
        return res

    def image_features(self, dets, img_off_host, imfeats, out):
        """Start features from the image (network.py:103-119, 223-240): boxes enlarged by
        half their size -> ROI max pooling of each image's feature map [1,H,W,C] (stride 16)
        -> flatten -> [FC imfeat_dim, relu ->] FC shortcut_dim, relu.  Returns
        (start_feat, roifeats[T,ph,pw,C], det_imfeats, frcn_boxes[T,5])."""
        T = dets.shape[0]
        if len(imfeats) != len(img_off_host) - 1:
            raise ValueError('one feature map per image expected (%d maps, %d images)'
                             % (len(imfeats), len(img_off_host) - 1))
        rois = ops.frcn_boxes(dets, 0.5, 0, out=self._buf('frcn_rois', (T, 5)))
        ph, pw_ = self.crop
        tops = []
        for i, fm in enumerate(imfeats):
            d0, d1 = int(img_off_host[i]), int(img_off_host[i + 1])
            top, _ = ops.roi_pool_fwd(fm, rois[d0:d1], ph, pw_, 1.0 / self.imfeat_stride)
            tops.append(top)
        roifeats = tops[0] if len(tops) == 1 else torch.cat(tops)
        x = roifeats.reshape(T, -1)
        scope = 'gnet/reduce_imfeats/fully_connected'
        if self.imfeat_dim > 0:
            x = self._fc(x, scope, True, out=self._buf('det_imfeats', (T, self.imfeat_dim)))
            scope += '_1'
        return self._fc(x, scope, True, out=out), roifeats, x, rois

    def forward(self, dets, scores, classes, img_off, imfeats=None, img_off_host=None,
                max_img=None):
        """dets[T,4] f32, scores[T] f32, classes[T] i32, img_off[B+1] i32 (all
        CUDA) -> dict(prediction[T], row_ptr, num_pairs, pair_c, pair_n,
        pair_iou, pw_feats, feats, [block_feats]).  Returned tensors are views
        of the engine's workspace: valid until the next forward.  With
        cfg.gnet.imfeats, `imfeats` is the list of per-image feature maps."""
        T = dets.shape[0]
        g = self.g
        if T == 0:
            return self._empty_result()
        if max_img is None and img_off_host is not None:
            max_img = int(np.max(np.diff(np.asarray(img_off_host))))
        row_ptr, num_pairs, pair_c, pair_n, pair_iou, cap = self.neighbors(dets, img_off, max_img)
        pw = self.pair_features(dets, scores, classes, pair_c, pair_n, pair_iou, num_pairs, cap)
        d = g['shortcut_dim']
        feats = self._buf('feats0', (T, d))
        im = None
        if self.imfeats:
            if imfeats is None:
                raise ValueError('cfg.gnet.imfeats is set: forward() needs the feature maps')
            if img_off_host is None:
                img_off_host = img_off.cpu().numpy()
            feats, *im = self.image_features(dets, img_off_host, imfeats, feats)
        else:
            feats.zero_()  # network.py:241-246
        block_feats = [feats.clone()] if self.keep_block_feats else None
        if self.fused_det and self.use_fused and self.use_tensor_cores and g['num_blocks'] > 0:
            feats = self._blocks_fused(feats, pair_c, pair_n, num_pairs, cap, pw, block_feats)
        else:
            for b in range(1, g['num_blocks'] + 1):
                out = self._buf('feats%d' % (b % 2), (T, d))
                feats = self.block(b, feats, row_ptr, pair_c, pair_n, num_pairs, cap, pw, out)
                if self.keep_block_feats:
                    block_feats.append(feats.clone())
        prediction = self.predict(feats)
        res = dict(prediction=prediction, row_ptr=row_ptr, num_pairs=num_pairs, pair_c=pair_c,
                   pair_n=pair_n, pair_iou=pair_iou, pw_feats=pw, feats=feats, capacity=cap)
        if self.keep_block_feats:
            res['block_feats'] = block_feats
        if im is not None:
            res['roifeats'], res['det_imfeats'], res['frcn_boxes'] = im
        return res

    # -------------------------------------------------------------- matching, loss
    def det_anno_iou(self, dets, classes, img_off_host, gt_boxes, gt_crowd, gt_classes,
                     gt_off_host):
        """Per-image dense [n_i, g_i] det/GT overlap blocks (network.py:174-187),
        concatenated; returns (iou_flat, iou_off_host[int64 B+1])."""
        B = len(img_off_host) - 1
        off = np.zeros(B + 1, dtype=np.int64)
        for i in range(B):
            n = img_off_host[i + 1] - img_off_host[i]
            gcount = gt_off_host[i + 1] - gt_off_host[i]
            off[i + 1] = off[i] + int(n) * int(gcount)
        flat = self._buf('det_anno_iou', (int(off[B]),))
        for i in range(B):
            d0, d1 = int(img_off_host[i]), int(img_off_host[i + 1])
            g0, g1 = int(gt_off_host[i]), int(gt_off_host[i + 1])
            if d1 == d0 or g1 == g0:
                continue
            out = flat[off[i]:off[i + 1]].view(1, d1 - d0, g1 - g0)
            ops.iou_dense(dets[d0:d1].unsqueeze(0), gt_boxes[g0:g1].unsqueeze(0),
                          crowd=gt_crowd[g0:g1].unsqueeze(0),
                          a_cls=classes[d0:d1].unsqueeze(0) if self.multiclass else None,
                          b_cls=gt_classes[g0:g1].unsqueeze(0) if self.multiclass else None,
                          out=out)
        return flat, off

    def matching_and_loss(self, prediction, dets, classes, img_off, img_off_host, gt_boxes,
                          gt_crowd, gt_classes, gt_off_host, class_weights, want_grad=False):
        """A9 + A10 (network.py:275-314) for the batch."""
        dev = self.device
        iou_flat, iou_off_host = self.det_anno_iou(dets, classes, img_off_host, gt_boxes, gt_crowd,
                                                   gt_classes, gt_off_host)
        iou_off = torch.from_numpy(iou_off_host).to(dev)
        gt_off = torch.from_numpy(np.asarray(gt_off_host, dtype=np.int32)).to(dev)
        max_gt = int(np.max(np.diff(gt_off_host))) if len(gt_off_host) > 1 else 0
        labels, weights, assignment = ops.detection_matching_batched(
            iou_flat, iou_off, prediction, gt_crowd, img_off, gt_off, max_gt)
        loss_out, dlogit = ops.loss_fwd(prediction, labels, weights, assignment, gt_crowd,
                                        gt_classes, img_off, gt_off, class_weights,
                                        self.normalize_loss, self.loss_multiplyer,
                                        want_grad=want_grad)
        return dict(labels=labels, weights=weights, det_gt_matching=assignment,
                    loss_out=loss_out, dlogit=dlogit, det_anno_iou=iou_flat,
                    det_anno_iou_off=iou_off_host)
