/*
 * gossipnet_b200 C ABI: the drop-in boundary of the B200-native GossipNet hot path.
 *
 * One shared library (gossipnet_b200/csrc/libgossipnet_b200.so), plain C linkage,
 * plain pointers and sizes, no torch / TensorFlow types.  The reference has no
 * C ABI of its own: its boundary is a TF-0.12 Python graph plus two
 * tf.load_op_library() objects.  Each entry point below names the reference
 * interface it replaces (paths relative to the reference repository root).
 * INTEGRATION.md shows the binding a maintainer of the reference would add.
 *
 * Conventions (all entry points):
 *   - every pointer is a DEVICE pointer on the current CUDA device unless the
 *     name ends in _host; the caller owns and pre-allocates every buffer
 *     (outputs and workspace); nothing is allocated or freed inside;
 *   - every call is asynchronous on `stream` (a cudaStream_t) and never
 *     synchronises; data-dependent sizes (the pair count P) stay on the device
 *     in an int32 counter the later calls read, so a whole forward pass is
 *     CUDA-graph capturable;
 *   - return value 0 = launched; non-zero = rejected before launch (argument
 *     validation, mirrors the reference ops' InvalidArgument checks) or a CUDA
 *     launch error; gn_last_error() returns a thread-local message;
 *   - re-entrant and thread-safe: no global mutable state besides the
 *     thread-local error string (the reference's DetectionMatchingOp keeps
 *     scratch vectors as kernel members, det_matching.cc:39-40, and is not);
 *   - detections of a batch of images are concatenated: dets[num_dets,4] with
 *     img_off[num_images+1] giving each image's [begin,end) rows.  All pair /
 *     neighbor indices are GLOBAL row numbers into that concatenation, so the
 *     block kernels never need to know about image boundaries.
 */
#ifndef GOSSIPNET_B200_H_
#define GOSSIPNET_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* gn_stream_t; /* cudaStream_t */

#define GN_OK 0
#define GN_ERR_INVALID_ARGUMENT 1
#define GN_ERR_CUDA 2
#define GN_ERR_UNSUPPORTED 3

const char* gn_last_error(void);
int gn_abi_version(void);
/* Number of SMs of the current device (grid sizing for persistent kernels). */
int gn_sm_count(void);

/* ---- A1 + A2: box IoU ------------------------------------------------------
 * Replaces Gnet._xyxy_to_boxdata / _intersection / _iou
 * (nms_net/network.py:462-472, :490-511, :474-488) and the multi-class
 * det<->GT mask (:177-187).
 *   a[batch,n,4], b[batch,m,4]  xyxy float32
 *   crowd[batch,m] (uint8, nullable): column j uses inter / area(a_i)
 *   a_cls[batch,n], b_cls[batch,m] (int32, both nullable): out=0 where classes differ
 *   out[batch,n,m] float32; bit-exact with the float32 reference arithmetic
 *   (no FMA contraction, IEEE division). */
int gn_iou_dense(const float* a, const float* b, const uint8_t* crowd,
                 const int32_t* a_cls, const int32_t* b_cls,
                 int batch, int n, int m, float* out, gn_stream_t stream);

/* Mask variant of the neighbor build (shipped path).  gn_neighbor_count_masks decides
 * iou >= thresh without the IEEE division except for borderline pairs (bit-identical decision,
 * see gn_neighbors.cu) and writes, next to degree[], one 32-bit hit mask per (row, 32 columns
 * of the row's image): masks[row * stride_words + w] covers columns img_lo + 32 w .. + 31;
 * stride_words >= ceil(largest image / 32).  gn_neighbor_fill_masks expands the masks into
 * pair_c / pair_n / pair_iou (exact IoU, computed for the hits only), same order and values
 * as gn_neighbor_fill. */
int gn_neighbor_count_masks(const float* dets, const int32_t* img_off, int num_images,
                            int num_dets, float thresh, int stride_words, int32_t* degree,
                            uint32_t* masks, gn_stream_t stream);
int gn_neighbor_fill_masks(const float* dets, const int32_t* img_off, int num_images,
                           int num_dets, const int32_t* row_ptr, int capacity,
                           const uint32_t* masks, int stride_words, int32_t* pair_c,
                           int32_t* pair_n, float* pair_iou, gn_stream_t stream);
/* ---- A3: neighbor build ----------------------------------------------------
 * Replaces tf.where(det_det_iou >= cfg.gnet.neighbor_thresh)
 * (nms_net/network.py:192-195).  IoU is recomputed from the boxes (same
 * arithmetic as gn_iou_dense), the dense N x N matrix is never materialised.
 * Pairs come out in the reference's row-major order: c ascending, then n
 * ascending; self pairs included.
 *
 *   gn_neighbor_count : degree[num_dets]  (pairs per row)
 *   gn_exclusive_scan : row_ptr[num_dets+1] from degree (row_ptr[num_dets] = P)
 *   gn_neighbor_fill  : pair_c[P], pair_n[P] (global rows, int32), pair_iou[P]
 *                       for pairs that fit `capacity`; *overflow is set to 1
 *                       (else left untouched) when P > capacity.
 * num_pairs for later calls is row_ptr + num_dets (device int32). */
int gn_neighbor_count(const float* dets, const int32_t* img_off, int num_images,
                      int num_dets, float thresh, int32_t* degree,
                      gn_stream_t stream);
int gn_exclusive_scan(const int32_t* in, int n, int32_t* out, gn_stream_t stream);
int gn_neighbor_fill(const float* dets, const int32_t* img_off, int num_images,
                     int num_dets, float thresh, const int32_t* row_ptr,
                     int capacity, int32_t* pair_c, int32_t* pair_n,
                     float* pair_iou, int32_t* overflow, gn_stream_t stream);

/* ---- A4: hand-crafted pair features ------------------------------------------
 * Replaces Gnet._geometry_feats (nms_net/network.py:411-454) times
 * cfg.gnet.pw_feat_multiplyer (:199-200).  out[capacity, width] with
 * width = 9 (num_classes<=1) or 2*num_classes+7; columns
 * [c_score | n_score | iou, x_dist, y_dist, l2_dist, w_diff, h_diff, aspect_diff].
 * Rows >= *num_pairs are not written. */
int gn_pair_geometry(const float* dets, const float* scores, const int32_t* classes,
                     const int32_t* pair_c, const int32_t* pair_n,
                     const float* pair_iou, const int32_t* num_pairs, int capacity,
                     int num_classes, float multiplier, float* out,
                     gn_stream_t stream);

/* ---- A5: pair-feature MLP, fused with A4 ------------------------------------
 * Replaces Gnet._pw_feats_fc (nms_net/network.py:324-342) applied to the
 * geometry features for the shipped shape num_pwfeat_fc=3:
 * width -> hidden -> hidden -> out_dim, ReLU after every layer.
 * w1[width,hidden] b1[hidden] w2[hidden,hidden] b2 w3[hidden,out_dim] b3;
 * pw_out[capacity,out_dim].  hidden must be 256 and out_dim 32 (the fused
 * kernel's shape); other shapes go through gn_pair_geometry + gn_fc_fwd.
 * The three layers run on the tensor cores (tcgen05, bf16 hi/lo split operands,
 * fp32 accumulation in TMEM).  wprep: caller-owned device workspace of
 * gn_pwfeat_prep_bytes() bytes (16-byte aligned) that receives the pre-split
 * operand image of the weights; it is rewritten by every call.
 * gn_pwfeat_mlp_fwd_ffma: same contract on the CUDA cores in fp32 (no wprep),
 * the in-library cross-check. */
int64_t gn_pwfeat_prep_bytes(void);
int gn_pwfeat_mlp_fwd(const float* dets, const float* scores, const int32_t* classes,
                      const int32_t* pair_c, const int32_t* pair_n,
                      const float* pair_iou, const int32_t* num_pairs, int capacity,
                      int num_classes, float multiplier,
                      const float* w1, const float* b1, const float* w2,
                      const float* b2, const float* w3, const float* b3,
                      int hidden, int out_dim, void* wprep, float* pw_out,
                      gn_stream_t stream);
int gn_pwfeat_mlp_fwd_ffma(const float* dets, const float* scores, const int32_t* classes,
                           const int32_t* pair_c, const int32_t* pair_n,
                           const float* pair_iou, const int32_t* num_pairs, int capacity,
                           int num_classes, float multiplier,
                           const float* w1, const float* b1, const float* w2,
                           const float* b2, const float* w3, const float* b3,
                           int hidden, int out_dim, float* pw_out, gn_stream_t stream);

/* ---- generic fully connected layer -----------------------------------------
 * tf.contrib.layers.fully_connected as the reference uses it everywhere:
 * y = act(x @ W[k,n] + b) (+ optional residual before the activation:
 * y = act(res + x @ W + b), the block shortcut network.py:407-408).
 * rows_dev (nullable) overrides `rows` with a device-side count (<= rows). */
int gn_fc_fwd(const float* x, int ldx, const float* w, const float* b,
              const float* residual, int ld_res, int relu, float* y, int ldy,
              int rows, const int32_t* rows_dev, int k, int n, gn_stream_t stream);

/* ---- A7: one Gnet block -----------------------------------------------------
 * Unfused pieces (reference formulation, nms_net/network.py:367-388):
 *   gn_block_gather_concat: x[P, w + 2r] = [pw | feats[pair_c] | nfeats[pair_n]],
 *                           n-part zeroed on self pairs (:368-376)
 *   gn_segment_max:         out[num_dets, f] = max over each row's pairs
 *                           (tf.segment_max, :387-388)
 * Fused hot path:
 *   gn_block_pair_fwd: gather + concat + pw_fc1 + pw_fc2 (+ReLU) + segment max
 *                      in one kernel, the two FCs on the tensor cores (tcgen05,
 *                      bf16 hi/lo split operands, fp32 accumulation in TMEM);
 *                      pooled[num_dets, f] must be zero-filled
 *                      by the caller (post-ReLU values are >= 0 and every row
 *                      has its self pair, so an integer atomicMax is exact). */
int gn_block_gather_concat(const float* pw, int w, const float* feats,
                           const float* nfeats, int r, const int32_t* pair_c,
                           const int32_t* pair_n, const int32_t* num_pairs,
                           int capacity, float* x, gn_stream_t stream);
int gn_segment_max(const float* x, int f, const int32_t* row_ptr, int num_dets,
                   float* out, gn_stream_t stream);
int gn_block_pair_fwd(const float* pw, int w, const float* feats,
                      const float* nfeats, int r, const int32_t* pair_c,
                      const int32_t* pair_n, const int32_t* num_pairs, int capacity,
                      const float* w1, const float* b1, const float* w2,
                      const float* b2, int f, float* pooled, gn_stream_t stream);
/* gn_block_pair_fwd_hl: same kernel, but feats / nfeats are the bf16 rows
 * [num_dets, 2r] = [r hi | r lo] that gn_block_det_fwd writes (hi = bf16(x),
 * lo = bf16(x - hi)): the operand chunks are copied, not re-split per pair. */
int gn_block_pair_fwd_hl(const float* pw, int w, const void* feats_hl,
                         const void* nfeats_hl, int r, const int32_t* pair_c,
                         const int32_t* pair_n, const int32_t* num_pairs, int capacity,
                         const float* w1, const float* b1, const float* w2,
                         const float* b2, const void* wimg, int f, float* pooled,
                         gn_stream_t stream);
/* Plain-bf16 arithmetic (BASELINE configs[2] names bf16): the same three fused tensor-core
 * kernels with bf16 operands (the hi parts of the operand images / activations only) and fp32
 * accumulation - one UMMA per k-step instead of three, half of the W2 stream.  Same
 * arguments as gn_pwfeat_mlp_fwd / gn_block_pair_fwd_pipe / gn_block_det_fwd_img.  Logits
 * agree with the fp32 path to ~1e-2 relative (tests state the tolerance); index outputs
 * (neighbor lists) do not depend on the mode. */
int gn_pwfeat_mlp_fwd_bf16(const float* dets, const float* scores, const int32_t* classes,
                           const int32_t* pair_c, const int32_t* pair_n, const float* pair_iou,
                           const int32_t* num_pairs, int capacity, int num_classes,
                           float multiplier, const float* w1, const float* b1, const float* w2,
                           const float* b2, const float* w3, const float* b3, int hidden,
                           int out_dim, void* wprep, float* pw_out, gn_stream_t stream);
int gn_block_pair_fwd_pipe_bf16(const float* pw, int w, const void* feats_hl,
                                const void* nfeats_hl, int r, const int32_t* pair_c,
                                const int32_t* pair_n, const int32_t* num_pairs, int capacity,
                                const float* b1, const float* b2, const void* wimg, int f,
                                float* pooled, gn_stream_t stream);
int gn_block_det_fwd_img_bf16(float* pooled, const float* feats_in, const void* wimg,
                              const float* b_fc1, const float* b_fc2, const float* b_rd,
                              int has_stage_a, int has_stage_b, float* feats_out, float* red_f32,
                              void* red_hl, const float* b_ab, float* ab_out, int num_dets,
                              int shortcut_dim, int pairfeat_dim, int reduced_dim,
                              gn_stream_t stream);
/* Predict head (A8, network.py:257-273): its hidden layers have activation_fn=None, so the
 * chain collapses to one affine map.  gn_predict_collapse folds the n_layers FCs described by
 * table (n_layers x 4 int32: weight offset, bias offset, in, out into flat_params; last out
 * must be 1) into w_eff[in_0] and b_eff[1] (scratch: 2 * max_dim floats); gn_rowdot_fwd
 * computes y[r] = x[r, :k] . w + b[0], one warp per row. */
int gn_predict_collapse(const float* flat_params, const int32_t* table, int n_layers,
                        int max_dim, float* scratch, float* w_eff, float* b_eff,
                        gn_stream_t stream);
int gn_rowdot_fwd(const float* x, int ldx, const float* w, const float* b, float* y, int rows,
                  int k, gn_stream_t stream);
/* gn_block_pair_fwd_pipe: the same pair stage (same operands as gn_block_pair_fwd_hl with
 * a prepared weight image, same results bit for bit) as a warp-specialised pipeline: one
 * persistent CTA per SM with fill / MMA / two epilogue warp groups and three tiles in
 * flight (csrc/gn_block_pipe.cu).  This is the variant the engine ships. */
int gn_block_pair_fwd_pipe(const float* pw, int w, const void* feats_hl,
                           const void* nfeats_hl, int r, const int32_t* pair_c,
                           const int32_t* pair_n, const int32_t* num_pairs, int capacity,
                           const float* b1, const float* b2, const void* wimg, int f,
                           float* pooled, gn_stream_t stream);
/* Prepared operand images.  Splitting / transposing the fp32 weights into the bf16
 * hi / lo K-major tiles inside every CTA of every launch is redundant work; with
 *   gn_prepare_operands(flat_params, table, entries, image)
 * ONE launch per forward converts all block weights (table: 6 int32 per matrix:
 * source offset in floats, k, n, byte offsets of the hi and lo tiles in `image`, chunk
 * pitch in bytes or 0 for n * 16), and
 * the kernels fetch their weights with a single bulk copy:
 *   wimg of gn_block_pair_fwd_hl (nullable; then w1 / w2 may be NULL):
 *        [pw_fc1^T hi | lo | pw_fc2^T hi | lo], gn_block_pair_image_bytes() bytes
 *   wimg of gn_block_det_fwd_img: [fc1^T hi | lo | fc2^T hi | lo | reduce_dim^T hi | lo |
 *        (W1[w:w+r] | W1[w+r:])^T hi | lo of the next block's pw_fc1],
 *        gn_block_det_image_bytes() bytes (parts of a skipped stage may be garbage).
 * A tile of a [k, n] weight holds chunk j (k = 8j..8j+7) of output column c at byte
 * offset j * n * 16 + c * 16. */
int64_t gn_block_pair_image_bytes(void);
int64_t gn_block_det_image_bytes(void);
int gn_prepare_operands(const float* flat_params, const int32_t* table, int entries,
                        void* image, gn_stream_t stream);
int gn_block_det_fwd_img(float* pooled, const float* feats_in, const void* wimg,
                         const float* b_fc1, const float* b_fc2, const float* b_rd,
                         int has_stage_a, int has_stage_b, float* feats_out,
                         float* red_f32, void* red_hl, const float* b_ab, float* ab_out,
                         int num_dets, int shortcut_dim, int pairfeat_dim, int reduced_dim,
                         gn_stream_t stream);
/* Pair stage with the first pair FC split by input block (x @ W1 = pw @ W1[0:w] +
 * feats[c] @ W1[w:w+r] + nfeats[n] @ W1[w+r:]): gn_block_det_fwd_img additionally
 * writes ab_out[num_dets, 2f] = [red @ W1[w:w+r] + b1 | red @ W1[w+r:]] (b_ab = b1 of the
 * NEXT block's pw_fc1; its weights are the 4th part of the det image), and
 *   gn_block_pair_fwd_ab: h1 = relu(pw @ W1[0:w] + ab[c, :f] + (c != n ? ab[n, f:] : 0)),
 *   then pw_fc2, ReLU and the segment max as gn_block_pair_fwd.
 * wimg: [W1[0:w]^T hi | lo | W2^T hi | lo], gn_block_pair_ab_image_bytes() bytes.
 * Same result as the reference formulation up to fp32 summation order. */
int64_t gn_block_pair_ab_image_bytes(void);
int gn_block_pair_fwd_ab(const float* pw, int w, const float* ab, int f,
                         const int32_t* pair_c, const int32_t* pair_n,
                         const int32_t* num_pairs, int capacity, const float* b2,
                         const void* wimg, float* pooled, gn_stream_t stream);
/* Detection-level layers fused across the block boundary (network.py:344-409), on
 * the tensor cores:
 *   stage A (pooled != NULL): d1 = relu(pooled @ w_fc1 + b_fc1);
 *           feats_out = relu(feats_in + d1 @ w_fc2 + b_fc2); pooled <- 0
 *   stage B (w_rd != NULL):   red = relu(X @ w_rd + b_rd), X = feats_out (or
 *           feats_in when stage A is skipped); written to red_f32[num_dets, r]
 *           and / or red_hl[num_dets, 2r] (bf16 hi | lo), either may be NULL.
 * Built for shortcut_dim 128, pairfeat_dim 64, reduced_dim 32, num_block_fc 2. */
int gn_block_det_fwd(float* pooled, const float* feats_in, const float* w_fc1,
                     const float* b_fc1, const float* w_fc2, const float* b_fc2,
                     const float* w_rd, const float* b_rd, float* feats_out,
                     float* red_f32, void* red_hl, int num_dets, int shortcut_dim,
                     int pairfeat_dim, int reduced_dim, gn_stream_t stream);
/* Same contract, evaluated with fp32 FFMA on the CUDA cores (no tensor cores):
 * the in-library cross-check of gn_block_pair_fwd's bf16x3 tensor-core product. */
int gn_block_pair_fwd_ffma(const float* pw, int w, const float* feats,
                           const float* nfeats, int r, const int32_t* pair_c,
                           const int32_t* pair_n, const int32_t* num_pairs, int capacity,
                           const float* w1, const float* b1, const float* w2,
                           const float* b2, int f, float* pooled, gn_stream_t stream);

/* ---- A9: DetectionMatching ---------------------------------------------------
 * Replaces the DetectionMatching TF op (nms_net/matching_module/det_matching.cc:
 * 16-33 op def, :72-160 Compute; python name matching_module.detection_matching,
 * __init__.py:13).  Batched over images: iou rows of image i are
 * [img_off[i], img_off[i+1]) and its GT columns [gt_off[i], gt_off[i+1]);
 * iou is stored per image as a dense [n_i, g_i] block at iou_off[i] (int64
 * element offsets).  Outputs are [num_dets].  Visiting orders reproduce the
 * reference's std::sort (+reverse) permutations exactly, including the order
 * among tied scores / equal ignore flags (libstdc++ introsort decision
 * sequence, gn_introsort.cuh), so results are bit-identical to a g++ build of
 * det_matching.cc for any non-NaN input.
 * max_gt: largest g_i of the batch (host-known; sizes the shared-memory GT
 * tables).  workspace: int32[gn_detection_matching_workspace_ints(num_dets)] (visiting
 * order + per-detection candidate records). */
int64_t gn_detection_matching_workspace_ints(int num_dets);
int gn_detection_matching(const float* iou, const int64_t* iou_off,
                          const float* score, const uint8_t* ignore,
                          const int32_t* img_off, const int32_t* gt_off,
                          int num_images, int num_dets, int max_gt,
                          float* labels, float* weights, int32_t* assignment,
                          int32_t* workspace, gn_stream_t stream);

/* ---- A10: loss ---------------------------------------------------------------
 * Replaces nms_net/network.py:281-313: class weighting of the matching
 * weights, sigmoid cross entropy, sum and mean.  gt_crowd / gt_classes are
 * concatenated like gt_off says; class_weights[num_classes+1].
 * weights_io is updated in place (weights * class_weights[det_class]);
 * loss_out[3*num_images] = per image (unnormed sum, normed mean, loss);
 * dlogit (nullable) [num_dets] = d loss / d prediction. */
int gn_loss_fwd(const float* prediction, const float* labels, float* weights_io,
                const int32_t* assignment, const uint8_t* gt_crowd,
                const int32_t* gt_classes, const int32_t* img_off,
                const int32_t* gt_off, int num_images, int num_dets,
                const float* class_weights, int normalize, float loss_multiplier,
                float* loss_out, float* dlogit, gn_stream_t stream);

/* clip_gradient_norm of slim.learning.create_train_op (reference train.py:73-76): every
 * parameter entry's gradient of the TOTAL loss, grad_scale * grads + decay * params (decay may be
 * NULL), is clipped by its own l2 norm (tf.clip_by_norm) and written back into grads.  table:
 * `entries` (offset, size) int32 pairs into the flat buffers.  Afterwards the optimizer steps run
 * with grad_scale = 1 and decay = NULL. */
int gn_clip_gradients(float* grads, const float* params, const float* decay, const int32_t* table,
                      int entries, float grad_scale, float clip_norm, gn_stream_t stream);

/* ---- A11: training step ----------------------------------------------------------
 * Backward pieces of the graph TF autodiff builds for nms_net/network.py and the
 * optimizer update of train.py:64-77.  The training forward runs the unfused
 * pieces above and keeps each FC's output, so every FC backward is
 *   gn_relu_mask     dy *= (y > 0)                      (tf.nn.relu grad, in place)
 *   gn_fc_bwd_weight dW[k,n] += x^T dy ; db[n] += colsum(dy)   (accumulates)
 *   dx = dy @ W^T    via gn_transpose + gn_fc_fwd(dy, W^T, zero bias)
 * rows_dev (nullable) overrides `rows` with a device-side count, as in gn_fc_fwd.
 *   gn_segment_max_bwd   tf.segment_max grad: rows equal to the max share the grad evenly
 *   gn_gather_concat_bwd grad of [pw | feats[c] | nfeats[n]] (network.py:367-376):
 *                        dpw_accum[P,w] += ; dfeats[T,r] += segment sums ;
 *                        dnfeats[T,r] += scatter (self pairs excluded)
 *   gn_add_inplace       dst += src
 *   gn_adam_step / gn_momentum_step  tf.train.AdamOptimizer / MomentumOptimizer on the
 *                        flat parameter buffer with g = grad_scale*grad + decay[i]*theta
 *                        (decay = weight_decay on regularised FC weights, else 0;
 *                        train.py:231, slim l2_regularizer); step counts from 1. */
int gn_relu_mask(float* dy, const float* y, int rows, const int32_t* rows_dev, int width,
                 gn_stream_t stream);
int gn_add_inplace(float* dst, const float* src, int rows, const int32_t* rows_dev, int width,
                   gn_stream_t stream);
int gn_transpose(const float* w, int k, int n, float* wt, gn_stream_t stream);
int gn_fc_bwd_weight(const float* x, int ldx, const float* dy, int ldy, float* dw, float* db,
                     int rows, const int32_t* rows_dev, int k, int n, gn_stream_t stream);
int gn_segment_max_bwd(const float* h, const float* pooled, const float* dpooled, int f,
                       const int32_t* row_ptr, int num_dets, float* dh, gn_stream_t stream);
int gn_gather_concat_bwd(const float* dx, int w, int r, const int32_t* pair_c,
                         const int32_t* pair_n, const int32_t* row_ptr, int num_dets,
                         const int32_t* num_pairs, int capacity, float* dpw_accum,
                         float* dfeats, float* dnfeats, gn_stream_t stream);
int gn_adam_step(float* params, const float* grads, float* m, float* v, const float* decay,
                 int64_t n, float lr, float beta1, float beta2, float eps, int64_t step,
                 float grad_scale, gn_stream_t stream);
int gn_momentum_step(float* params, const float* grads, float* accum, const float* decay,
                     int64_t n, float lr, float momentum, float grad_scale, gn_stream_t stream);

/* Tensor-core (tcgen05, bf16x3 = fp32 semantics) versions of gn_fc_fwd / gn_fc_bwd_weight for
 * the training step (tf.contrib.layers.fully_connected and its MatMul / BiasAdd / Relu
 * gradients, network.py:229-272, 328-341, 348-405):
 *   gn_prepare_fc_images  table: 6 int32 per entry (src offset in floats into flat_params, k, n,
 *                         dst byte offset into image, kpad, transposed).  transposed = 0: the
 *                         image of W[k,n] for y = x @ W (N = n, K = k padded to kpad);
 *                         transposed = 1: the image for dx = dy @ W^T (N = k, K = n padded).
 *                         Image bytes per entry: 2 * (kpad / 8) * N * 16.
 *   gn_fc_fwd_tc          y[rows,n] = act(residual + (x . (mask > 0)) @ W + b); mask (nullable)
 *                         has the shape and leading dimension of x and fuses the relu gradient
 *                         of the incoming dy; bias nullable (= 0).  k <= kpad <= 256, kpad % 16
 *                         == 0, n % 32 == 0, n <= 256, else GN_ERR_UNSUPPORTED.
 *   gn_fc_bwd_weight_tc   dW[k,n] += x^T @ (dy . (mask > 0)); db[n] += its column sums (fp32
 *                         atomics: the summation order is not fixed).  k <= 256,
 *                         n in {32, 64, 128, 256}. */
int gn_prepare_fc_images(const float* flat_params, const int32_t* table, int entries, void* image,
                         gn_stream_t stream);
int gn_fc_fwd_tc(const float* x, int ldx, const float* mask, const void* wimg, const float* bias,
                 const float* residual, int ld_res, int relu, float* y, int ldy, int rows,
                 const int32_t* rows_dev, int k, int kpad, int n, gn_stream_t stream);
int gn_fc_bwd_weight_tc(const float* x, int ldx, const float* dy, int ldy, const float* mask,
                        float* dw, float* db, int rows, const int32_t* rows_dev, int k, int n,
                        gn_stream_t stream);

/* gn_frcn_boxes: detection boxes -> rois of the image-feature head (network.py:78-100,
 * enlarge_windows + to_frcn_coords): rois[i] = (batch_index, cx - w (0.5 + padding),
 * cy - h (0.5 + padding), cx + w (0.5 + padding), cy + h (0.5 + padding)), float32 ops in the
 * reference's order (bit-exact). */
int gn_frcn_boxes(const float* dets, int num_dets, float padding, int batch_index, float* rois,
                  gn_stream_t stream);
/* ---- RoiPool / RoiPoolGrad ----------------------------------
 * Replace the RoiPool / RoiPoolGrad ops (nms_net/roi_pooling_layer/roi_pooling_op.cc:
 * 35-54 op definitions, :128-187 / :374-449 CPU semantics; python names
 * roi_pooling_op.roi_pool / roi_pool_grad, roi_pooling_op.py:6-7).
 *   bottom_data[batch,height,width,channels] f32 (NHWC), bottom_rois[num_rois,5] f32
 *   (batch_idx, x1, y1, x2, y2) -> top_data / argmax [num_rois,ph,pw,channels]
 *   (argmax: (h*W + w)*C + c inside the roi's image, -1 for an empty bin).
 *   gn_roi_pool_bwd: bottom_diff[batch,height,width,channels] = sum of top_diff over
 *   the pooled cells whose argmax is this element, accumulated in the reference's
 *   order (deterministic, bit-identical to the CPU op).
 * Bit-exact with the reference CPU kernels.  Errors: pooled sizes < 0 (:64-73). */
int gn_roi_pool_fwd(const float* bottom_data, int batch, int height, int width, int channels,
                    const float* bottom_rois, int num_rois, int pooled_height, int pooled_width,
                    float spatial_scale, float* top_data, int32_t* argmax, gn_stream_t stream);
int gn_roi_pool_bwd(int batch, int height, int width, int channels, const float* bottom_rois,
                    int num_rois, const int32_t* argmax, const float* top_diff,
                    int pooled_height, int pooled_width, float spatial_scale,
                    float* bottom_diff, gn_stream_t stream);

/* ---- diagnostics ---------------------------------------------------------------
 * c[128,64] = a[128,k] @ w[k,64] on the tensor cores with the building blocks of
 * the FC kernels (bf16x3 split operands, tcgen05.mma into TMEM, tcgen05.ld).
 * k: multiple of 16, <= 256.  No reference counterpart; used by the tests to pin
 * the descriptor / layout conventions independently of the fused kernels. */
int gn_selftest_umma(const float* a, const float* w, float* c, int k, gn_stream_t stream);
/* Same product with the A operand staged in tensor memory (tcgen05.st + the "TS" form of
 * tcgen05.mma): pins the A-in-TMEM layout. */
int gn_selftest_umma_ts(const float* a, const float* w, float* c, int k, gn_stream_t stream);
/* Micro-benchmark: `reps` back-to-back tcgen05.mma (M=128, N=n, K=16, bf16, SS) from
 * one thread per CTA, cycling over `distinct_b` B tiles; out_dev[0] = SM cycles,
 * out_dev[1] = reps (written by the last CTA to finish; all CTAs do the same work). */
int gn_selftest_umma_rate(int n, int reps, int distinct_b, int ctas, int64_t* out_dev,
                          gn_stream_t stream);

/* Tensor-map TMA conventions used by gn_block_pair_fwd_tma (no reference counterpart):
 * mat[rows,64] bf16, wmat[64,64] bf16, idx[128] int32 row indices.  dump[40960] = raw
 * shared-memory bytes of (tile load of rows row0..row0+127 | gather4 of rows idx | wmat)
 * in the SWIZZLE_128B layout; d_out[128,64] = (mat[row0:row0+128] + mat[idx]) @ wmat^T. */
int gn_selftest_tma(const void* mat_bf16, int rows, const void* wmat_bf16, const int32_t* idx,
                    int row0, void* dump, float* d_out, gn_stream_t stream);

/* gn_block_det_fwd_img writing only the detection-level half of the next block's pw_fc1
 * (network.py:376-386 split by input rows): u_out[num_dets,64] = red @ pw_fc1[32:64] + b_u,
 * next to red_hl.  Feeds gn_block_pair_fwd_tma.  plain_bf16 != 0: bf16 arithmetic. */
int gn_block_det_fwd_img_u(float* pooled, const float* feats_in, const void* wimg,
                           const float* b_fc1, const float* b_fc2, const float* b_rd,
                           int has_stage_a, int has_stage_b, float* feats_out, void* red_hl,
                           const float* b_u, float* u_out, int plain_bf16, int num_dets,
                           int shortcut_dim, int pairfeat_dim, int reduced_dim, gn_stream_t stream);

/* gn_block_det_fwd_img_u with stage A always present, on the copy-engine kernel (gn_det_tma.cu):
 * the shortcut tile comes in and the block output, u_out and red_hl leave by tensor-map TMA from a
 * dedicated warp, the pooled rows of the next tile are prefetched, so the eight epilogue warps
 * only run the four dependent GEMM epilogues (network.py:390-408, :348-354, :376-386).  Same
 * results bit for bit.  has_stage_b = 0: only feats_out (after the last block); pooled == NULL:
 * block 1, feats_in goes straight into reduce_dim and nothing is stored back. */
int gn_block_det_fwd_tma(float* pooled, const float* feats_in, const void* wimg,
                         const float* b_fc1, const float* b_fc2, const float* b_rd,
                         int has_stage_b, float* feats_out, void* red_hl, const float* b_u,
                         float* u_out, int plain_bf16, int num_dets, int shortcut_dim,
                         int pairfeat_dim, int reduced_dim, gn_stream_t stream);

/* Programmatic dependent launch of the persistent block kernels (gn_block_pair_fwd_tma,
 * gn_block_det_fwd_tma): each starts while its predecessor in the stream drains, runs its
 * prologue (barriers, tensor memory, weight image) and waits with griddepcontrol.wait before it
 * touches anything the predecessor wrote.  On by default; returns the previous setting. */
int gn_set_pdl(int enable);

/* Store-bandwidth micro-benchmark: writes `bytes` (multiple of 16384) bytes of constants with
 * mode 0 st.global.v4 | 1 st.global.cs.v4 | 2 st.global.v8 (256-bit) | 3 st.global.wt.v4 |
 * 4 cp.async.bulk shared->global (16 KB copies) | 5 st.global.v8 + L2 evict-first policy,
 * from sm_count * ctas_per_sm persistent CTAs.  The measured ceiling of the write-only dense
 * IoU kernel (bench.py roofline_iou.store_ceiling_gbs). */
int gn_selftest_store_bw(void* dst, int64_t bytes, int mode, int ctas_per_sm, gn_stream_t stream);

/* ---- A7 pair stage, TMA-fed (gn_block_tma.cu) ----------------------------------------
 * Same contract as gn_block_pair_fwd_pipe (network.py:367-388: gather/concat, pw_fc1,
 * pw_fc2, segment_max -> atomic max into pooled[T,64], which the caller zeroed), other
 * operand formats:
 *   pw_hl[capacity,64] bf16   pair features as (32 hi | 32 lo) rows (gn_pwfeat_mlp_fwd_hl)
 *   red_hl[num_dets+1,64] bf16 reduced features as (hi | lo) rows; row num_dets all zero
 *                             (the self pair's neighbor half, network.py:372-374)
 *   u[num_dets,u_pitch] f32   first 64 columns: red @ pw_fc1[32:64] + b_pw_fc1 (the
 *                             detection-level half of pw_fc1; gn_block_det_fwd_img ab_out)
 *   wimg                      gn_block_pair_tma_image_bytes() bytes of block b's image from
 *                             gn_prepare_pair_tma_image (table[b] = flat offsets of
 *                             pw_fc1/weights, pw_fc2/weights)
 * _bf16: plain bf16 operands (hi parts only), fp32 accumulation. */
int64_t gn_block_pair_tma_image_bytes(void);
int gn_prepare_pair_tma_image(const float* flat_params, const int32_t* table, int num_blocks,
                              void* image, gn_stream_t stream);
int gn_block_pair_fwd_tma(const void* pw_hl, const void* red_hl, int num_dets, const float* u,
                          int u_pitch, const int32_t* pair_c, const int32_t* pair_n,
                          const int32_t* num_pairs, int capacity, const float* b2,
                          const void* wimg, float* pooled, gn_stream_t stream);
int gn_block_pair_fwd_tma_bf16(const void* pw_hl, const void* red_hl, int num_dets, const float* u,
                               int u_pitch, const int32_t* pair_c, const int32_t* pair_n,
                               const int32_t* num_pairs, int capacity, const float* b2,
                               const void* wimg, float* pooled, gn_stream_t stream);
/* gn_pwfeat_mlp_fwd (network.py:324-342, 411-454) writing the pair features as bf16
 * (hi | lo) operand rows pw_hl[capacity,64]; pw_out (fp32 [capacity,32]) may be null.
 * plain_bf16 != 0: the bf16 arithmetic of gn_pwfeat_mlp_fwd_bf16. */
int gn_pwfeat_mlp_fwd_hl(const float* dets, const float* scores, const int32_t* classes,
                         const int32_t* pair_c, const int32_t* pair_n, const float* pair_iou,
                         const int32_t* num_pairs, int capacity, int num_classes,
                         float multiplier, const float* w1, const float* b1, const float* w2,
                         const float* b2, const float* w3, const float* b3, int hidden,
                         int out_dim, void* wprep, float* pw_out, void* pw_hl, int plain_bf16,
                         gn_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* GOSSIPNET_B200_H_ */
