"""Thin torch-tensor front-ends for the C ABI (include/gossipnet_b200.h).

torch is plumbing here: it owns device memory and the stream; every operation
below is one call into libgossipnet_b200.so on torch's current stream.  Inputs
must be CUDA tensors; there is no CPU path.
"""
import torch

from gossipnet_b200 import _lib


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _chk(t, dtype, name, allow_none=False):
    if t is None:
        if allow_none:
            return None
        raise ValueError('%s: tensor required' % name)
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise ValueError('%s must be a CUDA tensor (the hot path has no CPU fallback)' % name)
    if t.dtype != dtype:
        raise ValueError('%s must have dtype %s, got %s' % (name, dtype, t.dtype))
    if not t.is_contiguous():
        raise ValueError('%s must be contiguous' % name)
    return t.data_ptr()


# --------------------------------------------------------------------------- IoU
def iou_dense(a, b, crowd=None, a_cls=None, b_cls=None, out=None):
    """a[B,n,4] or [n,4]; b[B,m,4] or [m,4] -> out[B,n,m] / [n,m]."""
    squeeze = a.dim() == 2
    if squeeze:
        a, b = a.unsqueeze(0), b.unsqueeze(0)
        crowd = None if crowd is None else crowd.unsqueeze(0)
        a_cls = None if a_cls is None else a_cls.unsqueeze(0)
        b_cls = None if b_cls is None else b_cls.unsqueeze(0)
    if a.dim() != 3 or b.dim() != 3 or a.shape[-1] != 4 or b.shape[-1] != 4 \
            or a.shape[0] != b.shape[0]:
        raise ValueError('iou_dense expects a[B,n,4], b[B,m,4]')
    B, n, m = a.shape[0], a.shape[1], b.shape[1]
    if crowd is not None:
        if crowd.dtype == torch.bool:
            crowd = crowd.to(torch.uint8)
        if tuple(crowd.shape) != (B, m):
            raise ValueError('iou_dense: crowd must be [B,m]')
    if out is None:
        out = torch.empty((B, n, m), dtype=torch.float32, device=a.device)
    _lib.call('gn_iou_dense', _chk(a, torch.float32, 'a'), _chk(b, torch.float32, 'b'),
              _chk(crowd, torch.uint8, 'crowd', True), _chk(a_cls, torch.int32, 'a_cls', True),
              _chk(b_cls, torch.int32, 'b_cls', True), B, n, m,
              _chk(out, torch.float32, 'out'), _stream())
    return out[0] if squeeze else out


# --------------------------------------------------------------------- neighbors
def neighbor_count(dets, img_off, thresh, degree=None):
    n = dets.shape[0]
    if degree is None:
        degree = torch.empty(n, dtype=torch.int32, device=dets.device)
    _lib.call('gn_neighbor_count', _chk(dets, torch.float32, 'dets'),
              _chk(img_off, torch.int32, 'img_off'), img_off.numel() - 1, n, float(thresh),
              _chk(degree, torch.int32, 'degree'), _stream())
    return degree


def neighbor_count_masks(dets, img_off, thresh, stride_words, degree, masks):
    """Count pass of the mask variant: degree[T] + hit masks [T, stride_words] (uint32 as int32)."""
    _lib.call('gn_neighbor_count_masks', _chk(dets, torch.float32, 'dets'),
              _chk(img_off, torch.int32, 'img_off'), img_off.numel() - 1, dets.shape[0],
              float(thresh), int(stride_words), _chk(degree, torch.int32, 'degree'),
              _chk(masks, torch.int32, 'masks'), _stream())
    return degree, masks


def neighbor_fill_masks(dets, img_off, row_ptr, capacity, masks, stride_words, pair_c, pair_n,
                        pair_iou):
    _lib.call('gn_neighbor_fill_masks', _chk(dets, torch.float32, 'dets'),
              _chk(img_off, torch.int32, 'img_off'), img_off.numel() - 1, dets.shape[0],
              _chk(row_ptr, torch.int32, 'row_ptr'), int(capacity), _chk(masks, torch.int32, 'masks'),
              int(stride_words), _chk(pair_c, torch.int32, 'pair_c'),
              _chk(pair_n, torch.int32, 'pair_n'), _chk(pair_iou, torch.float32, 'pair_iou'),
              _stream())
    return pair_c, pair_n, pair_iou


def exclusive_scan(x, out=None):
    n = x.numel()
    if out is None:
        out = torch.empty(n + 1, dtype=torch.int32, device=x.device)
    _lib.call('gn_exclusive_scan', _chk(x, torch.int32, 'in'), n,
              _chk(out, torch.int32, 'out'), _stream())
    return out


def neighbor_fill(dets, img_off, thresh, row_ptr, capacity, pair_c, pair_n, pair_iou, overflow):
    _lib.call('gn_neighbor_fill', _chk(dets, torch.float32, 'dets'),
              _chk(img_off, torch.int32, 'img_off'), img_off.numel() - 1, dets.shape[0],
              float(thresh), _chk(row_ptr, torch.int32, 'row_ptr'), int(capacity),
              _chk(pair_c, torch.int32, 'pair_c'), _chk(pair_n, torch.int32, 'pair_n'),
              _chk(pair_iou, torch.float32, 'pair_iou'),
              _chk(overflow, torch.int32, 'overflow', True), _stream())


# ------------------------------------------------------------------ pair features
def pair_geometry(dets, scores, classes, pair_c, pair_n, pair_iou, num_pairs, capacity,
                  num_classes, multiplier, out=None):
    width = 9 if num_classes <= 1 else 2 * num_classes + 7
    if out is None:
        out = torch.empty((capacity, width), dtype=torch.float32, device=dets.device)
    _lib.call('gn_pair_geometry', _chk(dets, torch.float32, 'dets'),
              _chk(scores, torch.float32, 'scores'), _chk(classes, torch.int32, 'classes', True),
              _chk(pair_c, torch.int32, 'pair_c'), _chk(pair_n, torch.int32, 'pair_n'),
              _chk(pair_iou, torch.float32, 'pair_iou'), _chk(num_pairs, torch.int32, 'num_pairs'),
              int(capacity), int(num_classes), float(multiplier),
              _chk(out, torch.float32, 'out'), _stream())
    return out


_PREP = {}


def pwfeat_mlp_fwd(dets, scores, classes, pair_c, pair_n, pair_iou, num_pairs, capacity,
                   num_classes, multiplier, w1, b1, w2, b2, w3, b3, out=None, ffma=False,
                   wprep=None, bf16=False, out_hl=None, want_f32=True):
    """Fused geometry + 3-layer pair-feature MLP.  Tensor-core kernel by default
    (needs the `wprep` weight-image workspace; one per device is kept here when
    the caller does not pass its own); ffma=True runs the fp32 CUDA-core variant."""
    hidden, out_dim = w2.shape[1], w3.shape[1]
    if out is None and (want_f32 or out_hl is None):
        out = torch.empty((capacity, out_dim), dtype=torch.float32, device=dets.device)
    f32 = torch.float32
    args = [_chk(dets, f32, 'dets'), _chk(scores, f32, 'scores'),
            _chk(classes, torch.int32, 'classes', True), _chk(pair_c, torch.int32, 'pair_c'),
            _chk(pair_n, torch.int32, 'pair_n'), _chk(pair_iou, f32, 'pair_iou'),
            _chk(num_pairs, torch.int32, 'num_pairs'), int(capacity), int(num_classes),
            float(multiplier), _chk(w1, f32, 'w1'), _chk(b1, f32, 'b1'), _chk(w2, f32, 'w2'),
            _chk(b2, f32, 'b2'), _chk(w3, f32, 'w3'), _chk(b3, f32, 'b3'), int(hidden),
            int(out_dim)]
    if ffma:
        _lib.call('gn_pwfeat_mlp_fwd_ffma', *(args + [_chk(out, f32, 'pw_out'), _stream()]))
        return out
    if wprep is None:
        key = dets.device.index
        if key not in _PREP:
            _PREP[key] = torch.empty(int(_lib.load().gn_pwfeat_prep_bytes()), dtype=torch.uint8,
                                     device=dets.device)
        wprep = _PREP[key]
    if out_hl is not None:
        # bf16 (hi | lo) operand rows [capacity, 64] for gn_block_pair_fwd_tma; the fp32 rows
        # only when asked for
        _lib.call('gn_pwfeat_mlp_fwd_hl', *(args + [
            _chk(wprep, torch.uint8, 'wprep'),
            _chk(out, f32, 'pw_out') if (want_f32 and out is not None) else None,
            _chk(out_hl, torch.bfloat16, 'pw_hl'), 1 if bf16 else 0, _stream()]))
        return out if want_f32 else None
    _lib.call('gn_pwfeat_mlp_fwd_bf16' if bf16 else 'gn_pwfeat_mlp_fwd', *(args + [_chk(wprep, torch.uint8, 'wprep'),
                                             _chk(out, f32, 'pw_out'), _stream()]))
    return out


# ---------------------------------------------------------------------------- FC
def fc_fwd(x, w, b, relu, residual=None, out=None, rows_dev=None):
    """y = act(residual + x @ w + b); x[rows,k] (row stride may exceed k)."""
    if x.dim() != 2 or w.dim() != 2 or x.shape[1] != w.shape[0] or b.numel() != w.shape[1]:
        raise ValueError('fc_fwd: shape mismatch x%s w%s b%s' % (tuple(x.shape), tuple(w.shape),
                                                                 tuple(b.shape)))
    rows, k = x.shape
    n = w.shape[1]
    if out is None:
        out = torch.empty((rows, n), dtype=torch.float32, device=x.device)
    f32 = torch.float32
    _lib.call('gn_fc_fwd', _chk(x, f32, 'x'), x.stride(0), _chk(w, f32, 'w'), _chk(b, f32, 'b'),
              _chk(residual, f32, 'residual', True),
              0 if residual is None else residual.stride(0), 1 if relu else 0,
              _chk(out, f32, 'y'), out.stride(0), rows,
              _chk(rows_dev, torch.int32, 'rows_dev', True), k, n, _stream())
    return out


def prepare_fc_images(flat_params, table, image):
    """table[entries, 6] int32 (src offset, k, n, dst byte offset, kpad, transposed) -> bf16
    hi/lo K-major operand images of gn_fc_fwd_tc, one launch for all entries."""
    _lib.call('gn_prepare_fc_images', _chk(flat_params, torch.float32, 'flat_params'),
              _chk(table, torch.int32, 'table'), table.shape[0], _chk(image, torch.uint8, 'image'),
              _stream())


def fc_tc_supported(k, n):
    """Shapes gn_fc_fwd_tc / gn_fc_bwd_weight_tc take (everything the reference's FCs use
    except the 128 -> 1 logit layer)."""
    return k <= 256 and n <= 256 and n % 32 == 0


def fc_fwd_tc(x, wimg, k, n, bias, relu, residual=None, out=None, rows_dev=None, mask=None):
    """y = act(residual + (x * (mask > 0)) @ W + b) on the tensor cores (bf16x3); `wimg` is the
    operand image of W (gn_prepare_fc_images), k = true width of x, n = output width."""
    rows = x.shape[0]
    if x.dim() != 2 or x.shape[1] != k:
        raise ValueError('fc_fwd_tc: x must be [rows, %d]' % k)
    if mask is not None and tuple(mask.shape) != tuple(x.shape):
        raise ValueError('fc_fwd_tc: mask must have the shape of x')
    if out is None:
        out = torch.empty((rows, n), dtype=torch.float32, device=x.device)
    f32 = torch.float32
    kpad = (k + 15) // 16 * 16
    _lib.call('gn_fc_fwd_tc', _chk(x, f32, 'x'), k, _chk(mask, f32, 'mask', True),
              _chk(wimg, torch.uint8, 'wimg'), _chk(bias, f32, 'bias', True),
              _chk(residual, f32, 'residual', True), n if residual is not None else 0,
              1 if relu else 0, _chk(out, f32, 'y'), n, rows,
              _chk(rows_dev, torch.int32, 'rows_dev', True), k, kpad, n, _stream())
    return out


def fc_bwd_weight_tc(x, dy, dw, db, rows_dev=None, mask=None):
    """dW += x^T @ (dy * (mask > 0)); db += column sums of the masked dy (tensor cores)."""
    rows, k = x.shape
    n = dy.shape[1]
    f32 = torch.float32
    _lib.call('gn_fc_bwd_weight_tc', _chk(x, f32, 'x'), k, _chk(dy, f32, 'dy'), n,
              _chk(mask, f32, 'mask', True), _chk(dw, f32, 'dw'), _chk(db, f32, 'db', True), rows,
              _chk(rows_dev, torch.int32, 'rows_dev', True), k, n, _stream())


# ------------------------------------------------------------------------ blocks
def block_gather_concat(pw, feats, nfeats, pair_c, pair_n, num_pairs, capacity, out=None):
    w, r = pw.shape[1], feats.shape[1]
    if out is None:
        out = torch.empty((capacity, w + 2 * r), dtype=torch.float32, device=pw.device)
    f32 = torch.float32
    _lib.call('gn_block_gather_concat', _chk(pw, f32, 'pw'), w, _chk(feats, f32, 'feats'),
              _chk(nfeats, f32, 'nfeats'), r, _chk(pair_c, torch.int32, 'pair_c'),
              _chk(pair_n, torch.int32, 'pair_n'), _chk(num_pairs, torch.int32, 'num_pairs'),
              int(capacity), _chk(out, f32, 'x'), _stream())
    return out


def segment_max(x, row_ptr, num_dets, out=None):
    f = x.shape[1]
    if out is None:
        out = torch.empty((num_dets, f), dtype=torch.float32, device=x.device)
    _lib.call('gn_segment_max', _chk(x, torch.float32, 'x'), f,
              _chk(row_ptr, torch.int32, 'row_ptr'), int(num_dets),
              _chk(out, torch.float32, 'out'), _stream())
    return out


def block_pair_fwd(pw, feats, nfeats, pair_c, pair_n, num_pairs, capacity, w1, b1, w2, b2,
                   pooled, ffma=False, wimg=None):
    """pooled[num_dets,f] must be zero-filled; it is max-accumulated in place.
    ffma=True runs the fp32 CUDA-core variant instead of the tensor-core kernel."""
    f32 = torch.float32
    if feats.dtype == torch.bfloat16:
        # bf16 (hi | lo) rows from block_det_fwd: [num_dets, 2r]
        r = feats.shape[1] // 2
        _lib.call('gn_block_pair_fwd_hl', _chk(pw, f32, 'pw'), pw.shape[1],
                  _chk(feats, torch.bfloat16, 'feats_hl'), _chk(nfeats, torch.bfloat16, 'nfeats_hl'),
                  r, _chk(pair_c, torch.int32, 'pair_c'), _chk(pair_n, torch.int32, 'pair_n'),
                  _chk(num_pairs, torch.int32, 'num_pairs'), int(capacity),
                  _chk(w1, f32, 'w1', True), _chk(b1, f32, 'b1'), _chk(w2, f32, 'w2', True),
                  _chk(b2, f32, 'b2'), _chk(wimg, torch.uint8, 'wimg', True), b2.numel(),
                  _chk(pooled, f32, 'pooled'), _stream())
        return pooled
    _lib.call('gn_block_pair_fwd_ffma' if ffma else 'gn_block_pair_fwd', _chk(pw, f32, 'pw'), pw.shape[1], _chk(feats, f32, 'feats'),
              _chk(nfeats, f32, 'nfeats'), feats.shape[1], _chk(pair_c, torch.int32, 'pair_c'),
              _chk(pair_n, torch.int32, 'pair_n'), _chk(num_pairs, torch.int32, 'num_pairs'),
              int(capacity), _chk(w1, f32, 'w1'), _chk(b1, f32, 'b1'), _chk(w2, f32, 'w2'),
              _chk(b2, f32, 'b2'), w2.shape[1], _chk(pooled, f32, 'pooled'), _stream())
    return pooled


def block_det_fwd(pooled, feats_in, fc1, fc2, rd, feats_out=None, red_f32=None, red_hl=None):
    """Fused detection-level layers (see gn_block_det_fwd).  fc1 / fc2 / rd are
    (weights, biases) pairs or None; pooled=None skips stage A, rd=None stage B."""
    f32 = torch.float32
    T, d = feats_in.shape
    f = fc1[0].shape[0] if fc1 is not None else 64
    r = rd[0].shape[1] if rd is not None else 32
    nul = (None, None)
    w1, b1 = fc1 if fc1 is not None else nul
    w2, b2 = fc2 if fc2 is not None else nul
    wr, br = rd if rd is not None else nul
    _lib.call('gn_block_det_fwd', _chk(pooled, f32, 'pooled', True), _chk(feats_in, f32, 'feats_in'),
              _chk(w1, f32, 'w_fc1', True), _chk(b1, f32, 'b_fc1', True),
              _chk(w2, f32, 'w_fc2', True), _chk(b2, f32, 'b_fc2', True),
              _chk(wr, f32, 'w_rd', True), _chk(br, f32, 'b_rd', True),
              _chk(feats_out, f32, 'feats_out', True), _chk(red_f32, f32, 'red_f32', True),
              _chk(red_hl, torch.bfloat16, 'red_hl', True), T, d, f, r, _stream())


def prepare_operands(flat_params, table, image):
    """One launch: every [k,n] weight listed in `table` (int32 [entries,5]) -> bf16 hi/lo
    K-major operand tiles inside `image` (uint8)."""
    _lib.call('gn_prepare_operands', _chk(flat_params, torch.float32, 'flat_params'),
              _chk(table, torch.int32, 'table'), table.shape[0],
              _chk(image, torch.uint8, 'image'), _stream())


def block_det_fwd_img(pooled, feats_in, wimg, b_fc1, b_fc2, b_rd, feats_out=None, red_f32=None,
                      red_hl=None, b_ab=None, ab_out=None, bf16=False):
    """gn_block_det_fwd with weights from a prepared operand image; a stage runs when
    its bias is given (b_fc1 & b_fc2 -> stage A, b_rd -> stage B); ab_out (+ b_ab)
    additionally requests the per-detection halves of the next pair FC."""
    f32 = torch.float32
    T, d = feats_in.shape
    _lib.call('gn_block_det_fwd_img_bf16' if bf16 else 'gn_block_det_fwd_img',
              _chk(pooled, f32, 'pooled', True),
              _chk(feats_in, f32, 'feats_in'), _chk(wimg, torch.uint8, 'wimg'),
              _chk(b_fc1, f32, 'b_fc1', True), _chk(b_fc2, f32, 'b_fc2', True),
              _chk(b_rd, f32, 'b_rd', True), 1 if b_fc1 is not None else 0,
              1 if b_rd is not None else 0, _chk(feats_out, f32, 'feats_out', True),
              _chk(red_f32, f32, 'red_f32', True), _chk(red_hl, torch.bfloat16, 'red_hl', True),
              _chk(b_ab, f32, 'b_ab', True), _chk(ab_out, f32, 'ab_out', True),
              T, d, 64, 32, _stream())


def block_det_fwd_img_u(pooled, feats_in, wimg, b_fc1, b_fc2, b_rd, feats_out, red_hl, b_u, u_out,
                        bf16=False):
    """block_det_fwd_img for the TMA-fed pair stage: next to red_hl it writes only
    u_out[T,64] = red @ pw_fc1[32:64] + b_u (the detection-level term of the next pw_fc1)."""
    f32 = torch.float32
    T, d = feats_in.shape
    _lib.call('gn_block_det_fwd_img_u', _chk(pooled, f32, 'pooled', True),
              _chk(feats_in, f32, 'feats_in'), _chk(wimg, torch.uint8, 'wimg'),
              _chk(b_fc1, f32, 'b_fc1', True), _chk(b_fc2, f32, 'b_fc2', True),
              _chk(b_rd, f32, 'b_rd', True), 1 if b_fc1 is not None else 0,
              1 if b_rd is not None else 0, _chk(feats_out, f32, 'feats_out', True),
              _chk(red_hl, torch.bfloat16, 'red_hl', True), _chk(b_u, f32, 'b_u', True),
              _chk(u_out, f32, 'u_out', True), 1 if bf16 else 0, T, d, 64, 32, _stream())


def block_det_fwd_tma(pooled, feats_in, wimg, b_fc1, b_fc2, b_rd, feats_out, red_hl, b_u, u_out,
                      bf16=False):
    """block_det_fwd_img_u on the copy-engine kernel: tile transfers by tensor-map TMA from a
    dedicated warp (gn_det_tma.cu).  b_rd None: only feats_out (after the last block);
    pooled None: block 1, feats_in straight into reduce_dim."""
    f32 = torch.float32
    T, d = feats_in.shape
    _lib.call('gn_block_det_fwd_tma', _chk(pooled, f32, 'pooled', True), _chk(feats_in, f32, 'feats_in'),
              _chk(wimg, torch.uint8, 'wimg'), _chk(b_fc1, f32, 'b_fc1', True),
              _chk(b_fc2, f32, 'b_fc2', True), _chk(b_rd, f32, 'b_rd', True),
              1 if b_rd is not None else 0, _chk(feats_out, f32, 'feats_out', True),
              _chk(red_hl, torch.bfloat16, 'red_hl', True),
              _chk(b_u, f32, 'b_u', True), _chk(u_out, f32, 'u_out', True), 1 if bf16 else 0, T, d,
              64, 32, _stream())


def set_pdl(enable):
    """Programmatic dependent launch of the persistent block kernels (gn_set_pdl); returns the
    previous setting.  A captured CUDA graph keeps the setting it was captured with."""
    return bool(_lib.load().gn_set_pdl(1 if enable else 0))


def predict_collapse(flat, table, max_dim, scratch, w_eff, b_eff):
    """Fold the linear predict head into (w_eff, b_eff) (see gn_predict_collapse)."""
    f32 = torch.float32
    _lib.call('gn_predict_collapse', _chk(flat, f32, 'flat'), _chk(table, torch.int32, 'table'),
              table.shape[0], int(max_dim), _chk(scratch, f32, 'scratch'), _chk(w_eff, f32, 'w_eff'),
              _chk(b_eff, f32, 'b_eff'), _stream())


def rowdot_fwd(x, w, b, out):
    """out[r] = x[r] . w + b[0]."""
    f32 = torch.float32
    _lib.call('gn_rowdot_fwd', _chk(x, f32, 'x'), x.stride(0), _chk(w, f32, 'w'), _chk(b, f32, 'b'),
              _chk(out, f32, 'out'), x.shape[0], x.shape[1], _stream())
    return out


def block_pair_fwd_pipe(pw, feats_hl, pair_c, pair_n, num_pairs, capacity, b1, b2, wimg, pooled,
                        bf16=False):
    """Pipelined pair stage (see gn_block_pair_fwd_pipe): bf16 (hi | lo) reduced-feature
    rows [num_dets, 2r], prepared weight image; pooled must be zero-filled."""
    f32 = torch.float32
    _lib.call('gn_block_pair_fwd_pipe_bf16' if bf16 else 'gn_block_pair_fwd_pipe',
              _chk(pw, f32, 'pw'), pw.shape[1],
              _chk(feats_hl, torch.bfloat16, 'feats_hl'), _chk(feats_hl, torch.bfloat16, 'nfeats_hl'),
              feats_hl.shape[1] // 2, _chk(pair_c, torch.int32, 'pair_c'),
              _chk(pair_n, torch.int32, 'pair_n'), _chk(num_pairs, torch.int32, 'num_pairs'),
              int(capacity), _chk(b1, f32, 'b1'), _chk(b2, f32, 'b2'),
              _chk(wimg, torch.uint8, 'wimg'), b2.numel(), _chk(pooled, f32, 'pooled'), _stream())
    return pooled


def pair_tma_image_bytes():
    return int(_lib.load().gn_block_pair_tma_image_bytes())


def prepare_pair_tma_image(flat_params, table, image):
    """table[num_blocks, 2] int32 (flat offsets of pw_fc1 / pw_fc2 weights) -> the swizzled
    operand images of gn_block_pair_fwd_tma, one launch for all blocks."""
    _lib.call('gn_prepare_pair_tma_image', _chk(flat_params, torch.float32, 'flat_params'),
              _chk(table, torch.int32, 'table'), table.shape[0], _chk(image, torch.uint8, 'image'),
              _stream())


def block_pair_fwd_tma(pw_hl, red_hl, num_dets, u, pair_c, pair_n, num_pairs, capacity, b2, wimg,
                       pooled, bf16=False):
    """Pair stage with TMA-fed operands: pw_hl[capacity,64] bf16 (hi|lo) rows, red_hl[T+1,64]
    bf16 (hi|lo) reduced features with an all-zero row T, u[T,pitch] fp32 = red @ W1[32:64] + b1
    in its first 64 columns."""
    if red_hl.shape[0] < num_dets + 1:
        raise ValueError('red_hl needs %d rows (zero row at the end)' % (num_dets + 1))
    _lib.call('gn_block_pair_fwd_tma_bf16' if bf16 else 'gn_block_pair_fwd_tma',
              _chk(pw_hl, torch.bfloat16, 'pw_hl'), _chk(red_hl, torch.bfloat16, 'red_hl'),
              int(num_dets), _chk(u, torch.float32, 'u'), int(u.shape[1]),
              _chk(pair_c, torch.int32, 'pair_c'), _chk(pair_n, torch.int32, 'pair_n'),
              _chk(num_pairs, torch.int32, 'num_pairs'), int(capacity),
              _chk(b2, torch.float32, 'b2'), _chk(wimg, torch.uint8, 'wimg'),
              _chk(pooled, torch.float32, 'pooled'), _stream())
    return pooled


def block_pair_fwd_ab(pw, ab, pair_c, pair_n, num_pairs, capacity, b2, wimg, pooled):
    """Pair stage on per-detection halves AB[T, 2f] (see gn_block_pair_fwd_ab)."""
    f32 = torch.float32
    _lib.call('gn_block_pair_fwd_ab', _chk(pw, f32, 'pw'), pw.shape[1], _chk(ab, f32, 'ab'),
              ab.shape[1] // 2, _chk(pair_c, torch.int32, 'pair_c'),
              _chk(pair_n, torch.int32, 'pair_n'), _chk(num_pairs, torch.int32, 'num_pairs'),
              int(capacity), _chk(b2, f32, 'b2'), _chk(wimg, torch.uint8, 'wimg'),
              _chk(pooled, f32, 'pooled'), _stream())
    return pooled


# ---------------------------------------------------------------- matching, loss
def detection_matching_batched(iou, iou_off, score, ignore, img_off, gt_off, max_gt):
    n = score.numel()
    dev = score.device
    labels = torch.empty(n, dtype=torch.float32, device=dev)
    weights = torch.empty(n, dtype=torch.float32, device=dev)
    assignment = torch.empty(n, dtype=torch.int32, device=dev)
    ws = torch.empty(int(_lib.load().gn_detection_matching_workspace_ints(n)), dtype=torch.int32,
                     device=dev)
    _lib.call('gn_detection_matching', _chk(iou, torch.float32, 'iou'),
              _chk(iou_off, torch.int64, 'iou_off'), _chk(score, torch.float32, 'score'),
              _chk(ignore, torch.uint8, 'ignore'), _chk(img_off, torch.int32, 'img_off'),
              _chk(gt_off, torch.int32, 'gt_off'), img_off.numel() - 1, n, int(max_gt),
              _chk(labels, torch.float32, 'labels'), _chk(weights, torch.float32, 'weights'),
              _chk(assignment, torch.int32, 'assignment'), _chk(ws, torch.int32, 'workspace'),
              _stream())
    return labels, weights, assignment


def loss_fwd(prediction, labels, weights, assignment, gt_crowd, gt_classes, img_off, gt_off,
             class_weights, normalize, loss_multiplier, want_grad=False):
    """Updates `weights` in place; returns (loss_out[num_images,3], dlogit|None)."""
    num_images = img_off.numel() - 1
    dev = prediction.device
    loss_out = torch.empty((num_images, 3), dtype=torch.float32, device=dev)
    dlogit = torch.empty_like(prediction) if want_grad else None
    f32 = torch.float32
    _lib.call('gn_loss_fwd', _chk(prediction, f32, 'prediction'), _chk(labels, f32, 'labels'),
              _chk(weights, f32, 'weights'), _chk(assignment, torch.int32, 'assignment'),
              _chk(gt_crowd, torch.uint8, 'gt_crowd'), _chk(gt_classes, torch.int32, 'gt_classes'),
              _chk(img_off, torch.int32, 'img_off'), _chk(gt_off, torch.int32, 'gt_off'),
              num_images, prediction.numel(), _chk(class_weights, f32, 'class_weights'),
              1 if normalize else 0, float(loss_multiplier), _chk(loss_out, f32, 'loss_out'),
              _chk(dlogit, f32, 'dlogit', True), _stream())
    return loss_out, dlogit


# ---------------------------------------------------------------------- training
def relu_mask(dy, y, rows_dev=None):
    """dy *= (y > 0) in place (rows beyond *rows_dev untouched)."""
    rows, width = dy.shape
    _lib.call('gn_relu_mask', _chk(dy, torch.float32, 'dy'), _chk(y, torch.float32, 'y'), rows,
              _chk(rows_dev, torch.int32, 'rows_dev', True), width, _stream())
    return dy


def add_inplace(dst, src, rows_dev=None):
    rows, width = (dst.shape[0], dst.numel() // max(dst.shape[0], 1)) if dst.dim() > 1 \
        else (1, dst.numel())
    _lib.call('gn_add_inplace', _chk(dst, torch.float32, 'dst'), _chk(src, torch.float32, 'src'),
              rows, _chk(rows_dev, torch.int32, 'rows_dev', True), width, _stream())
    return dst


def transpose(w, out=None):
    k, n = w.shape
    if out is None:
        out = torch.empty((n, k), dtype=torch.float32, device=w.device)
    _lib.call('gn_transpose', _chk(w, torch.float32, 'w'), k, n, _chk(out, torch.float32, 'wt'),
              _stream())
    return out


def fc_bwd_weight(x, dy, dw, db, rows_dev=None):
    """dw[k,n] += x^T dy; db[n] += colsum(dy)."""
    rows, k = x.shape
    n = dy.shape[1]
    _lib.call('gn_fc_bwd_weight', _chk(x, torch.float32, 'x'), x.stride(0),
              _chk(dy, torch.float32, 'dy'), dy.stride(0), _chk(dw, torch.float32, 'dw'),
              _chk(db, torch.float32, 'db', True), rows,
              _chk(rows_dev, torch.int32, 'rows_dev', True), k, n, _stream())


def segment_max_bwd(h, pooled, dpooled, row_ptr, out):
    num_dets, f = pooled.shape
    _lib.call('gn_segment_max_bwd', _chk(h, torch.float32, 'h'),
              _chk(pooled, torch.float32, 'pooled'), _chk(dpooled, torch.float32, 'dpooled'), f,
              _chk(row_ptr, torch.int32, 'row_ptr'), num_dets, _chk(out, torch.float32, 'dh'),
              _stream())
    return out


def gather_concat_bwd(dx, w, r, pair_c, pair_n, row_ptr, num_dets, num_pairs, capacity,
                      dpw_accum, dfeats, dnfeats):
    _lib.call('gn_gather_concat_bwd', _chk(dx, torch.float32, 'dx'), w, r,
              _chk(pair_c, torch.int32, 'pair_c'), _chk(pair_n, torch.int32, 'pair_n'),
              _chk(row_ptr, torch.int32, 'row_ptr'), num_dets,
              _chk(num_pairs, torch.int32, 'num_pairs'), int(capacity),
              _chk(dpw_accum, torch.float32, 'dpw_accum'), _chk(dfeats, torch.float32, 'dfeats'),
              _chk(dnfeats, torch.float32, 'dnfeats'), _stream())


def adam_step(params, grads, m, v, decay, lr, beta1, beta2, eps, step, grad_scale):
    _lib.call('gn_adam_step', _chk(params, torch.float32, 'params'),
              _chk(grads, torch.float32, 'grads'), _chk(m, torch.float32, 'm'),
              _chk(v, torch.float32, 'v'), _chk(decay, torch.float32, 'decay', True),
              params.numel(), float(lr), float(beta1), float(beta2), float(eps), int(step),
              float(grad_scale), _stream())


def clip_gradients(grads, params, decay, table, grad_scale, clip_norm):
    """grads <- clip_by_norm(grad_scale * grads + decay * params) per parameter entry
    (table: int32 [entries, 2] = offset, size), the reference's clip_gradient_norm."""
    _lib.call('gn_clip_gradients', _chk(grads, torch.float32, 'grads'),
              _chk(params, torch.float32, 'params'), _chk(decay, torch.float32, 'decay', True),
              _chk(table, torch.int32, 'table'), table.shape[0], float(grad_scale),
              float(clip_norm), _stream())


def momentum_step(params, grads, accum, decay, lr, momentum, grad_scale):
    _lib.call('gn_momentum_step', _chk(params, torch.float32, 'params'),
              _chk(grads, torch.float32, 'grads'), _chk(accum, torch.float32, 'accum'),
              _chk(decay, torch.float32, 'decay', True), params.numel(), float(lr),
              float(momentum), float(grad_scale), _stream())


# ---------------------------------------------------------------------- roi pool
def _roi_args(bottom_data, bottom_rois):
    # the reference op's rank checks (roi_pooling_op.cc:88-93)
    if bottom_data.dim() != 4:
        raise ValueError('data must be 4-dimensional')
    if bottom_rois.dim() != 2:
        raise ValueError('rois must be 2-dimensional')
    if bottom_rois.shape[1] != 5:
        raise ValueError('rois must be [num_rois, 5] (batch_idx, x1, y1, x2, y2)')


def frcn_boxes(dets, padding=0.5, batch_index=0, out=None):
    """dets[N,4] -> rois[N,5] = (batch_index, enlarged box) (network.py:78-100)."""
    n = dets.shape[0]
    if out is None:
        out = torch.empty((n, 5), dtype=torch.float32, device=dets.device)
    _lib.call('gn_frcn_boxes', _chk(dets, torch.float32, 'dets'), n, float(padding), int(batch_index),
              _chk(out, torch.float32, 'rois'), _stream())
    return out


def roi_pool_fwd(bottom_data, bottom_rois, pooled_height, pooled_width, spatial_scale):
    _roi_args(bottom_data, bottom_rois)
    if pooled_height < 0:      # roi_pooling_op.cc:64-73
        raise ValueError('Need pooled_height >= 0, got %d' % pooled_height)
    if pooled_width < 0:
        raise ValueError('Need pooled_width >= 0, got %d' % pooled_width)
    b, h, w, c = bottom_data.shape
    r = bottom_rois.shape[0]
    shape = (r, int(pooled_height), int(pooled_width), c)
    top = torch.empty(shape, dtype=torch.float32, device=bottom_data.device)
    argmax = torch.empty(shape, dtype=torch.int32, device=bottom_data.device)
    _lib.call('gn_roi_pool_fwd', _chk(bottom_data, torch.float32, 'bottom_data'), b, h, w, c,
              _chk(bottom_rois, torch.float32, 'bottom_rois'), r, int(pooled_height),
              int(pooled_width), float(spatial_scale), _chk(top, torch.float32, 'top_data'),
              _chk(argmax, torch.int32, 'argmax'), _stream())
    return top, argmax


def roi_pool_bwd(bottom_data, bottom_rois, argmax, grad, pooled_height, pooled_width,
                 spatial_scale):
    _roi_args(bottom_data, bottom_rois)
    if argmax.dim() != 4:
        raise ValueError('argmax_data must be 4-dimensional')
    if grad.dim() != 4:
        raise ValueError('out_backprop must be 4-dimensional')
    b, h, w, c = bottom_data.shape
    out = torch.empty_like(bottom_data)
    _lib.call('gn_roi_pool_bwd', b, h, w, c, _chk(bottom_rois, torch.float32, 'bottom_rois'),
              bottom_rois.shape[0], _chk(argmax, torch.int32, 'argmax'),
              _chk(grad, torch.float32, 'grad'), int(pooled_height), int(pooled_width),
              float(spatial_scale), _chk(out, torch.float32, 'bottom_diff'), _stream())
    return out


# -------------------------------------------------------------------- diagnostics
def selftest_umma(a, w, a_in_tmem=False):
    """c[128,64] = a[128,k] @ w[k,64] through tcgen05 (see gn_selftest.cu); a_in_tmem
    stages A in tensor memory (tcgen05.st + TS-form UMMA) instead of shared memory."""
    if tuple(a.shape)[0] != 128 or tuple(w.shape) != (a.shape[1], 64):
        raise ValueError('selftest_umma expects a[128,k], w[k,64]')
    c = torch.empty((128, 64), dtype=torch.float32, device=a.device)
    _lib.call('gn_selftest_umma_ts' if a_in_tmem else 'gn_selftest_umma', _chk(a, torch.float32, 'a'), _chk(w, torch.float32, 'w'),
              _chk(c, torch.float32, 'c'), int(a.shape[1]), _stream())
    return c


def selftest_tma(mat, wmat, idx, row0):
    """Tensor-map TMA conventions (gn_selftest.cu): mat[rows,64] bf16, wmat[64,64] bf16,
    idx[128] int32 -> (raw dump of the A0 | A1 | B shared-memory tiles as uint8[40960],
    D[128,64] = (mat[row0:row0+128] + mat[idx]) @ wmat^T)."""
    dump = torch.zeros(2 * 16384 + 8192, dtype=torch.uint8, device=mat.device)
    d = torch.zeros((128, 64), dtype=torch.float32, device=mat.device)
    _lib.call('gn_selftest_tma', _chk(mat, torch.bfloat16, 'mat'), mat.shape[0],
              _chk(wmat, torch.bfloat16, 'wmat'), _chk(idx, torch.int32, 'idx'), int(row0),
              _chk(dump, torch.uint8, 'dump'), _chk(d, torch.float32, 'd'), _stream())
    return dump, d


STORE_BW_MODES = ('st.v4', 'st.cs.v4', 'st.v8', 'st.wt.v4', 'bulk_s2g_16k', 'st.v8.evict_first')


def selftest_store_bw(buf, mode, ctas_per_sm=8):
    """Fill `buf` (uint8, multiple of 16 KiB) with the library's store micro-benchmark;
    mode: a name from STORE_BW_MODES."""
    n = buf.numel() // 16384 * 16384
    _lib.call('gn_selftest_store_bw', _chk(buf, torch.uint8, 'buf'), n,
              STORE_BW_MODES.index(mode), int(ctas_per_sm), _stream())
