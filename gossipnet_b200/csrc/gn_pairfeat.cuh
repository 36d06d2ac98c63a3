// Per-pair geometry features (A4), shared by the raw-feature kernel and the
// fused pair-feature MLP.  Operation order and rounding follow
// nms_net/network.py:428-450 (every op rounded to float32, no contraction).
#pragma once
#include "gn_common.cuh"

namespace gn {

// ln(2) rounded to float32, i.e. np.float32(np.log(2.0)) (network.py:447)
#define GN_LN2_F32 0.693147182464599609375f

// g[0..6] = iou, x_dist, y_dist, l2_dist, w_diff, h_diff, aspect_diff
__device__ __forceinline__ void pair_geometry(float4 cb, float4 nb, float iou, float mult,
                                              float* g) {
  const float c_w = __fsub_rn(cb.z, cb.x), c_h = __fsub_rn(cb.w, cb.y);
  const float n_w = __fsub_rn(nb.z, nb.x), n_h = __fsub_rn(nb.w, nb.y);
  const float c_scale = __fdiv_rn(__fadd_rn(c_w, c_h), 2.0f);
  const float c_cx = __fadd_rn(cb.x, __fdiv_rn(c_w, 2.0f));
  const float c_cy = __fadd_rn(cb.y, __fdiv_rn(c_h, 2.0f));
  const float n_cx = __fadd_rn(nb.x, __fdiv_rn(n_w, 2.0f));
  const float n_cy = __fadd_rn(nb.y, __fdiv_rn(n_h, 2.0f));
  const float dx = __fsub_rn(n_cx, c_cx), dy = __fsub_rn(n_cy, c_cy);
  const float l2 = __fdiv_rn(__fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy))), c_scale);
  const float w_diff = __fdiv_rn(logf(__fdiv_rn(n_w, c_w)), GN_LN2_F32);
  const float h_diff = __fdiv_rn(logf(__fdiv_rn(n_h, c_h)), GN_LN2_F32);
  const float asp = __fdiv_rn(
      __fsub_rn(logf(__fdiv_rn(n_w, n_h)), logf(__fdiv_rn(c_w, c_h))), GN_LN2_F32);
  g[0] = __fmul_rn(iou, mult);
  g[1] = __fmul_rn(__fdiv_rn(dx, c_scale), mult);
  g[2] = __fmul_rn(__fdiv_rn(dy, c_scale), mult);
  g[3] = __fmul_rn(l2, mult);
  g[4] = __fmul_rn(w_diff, mult);
  g[5] = __fmul_rn(h_diff, mult);
  g[6] = __fmul_rn(asp, mult);
}

// The same features in two independent halves (identical arithmetic), so that two
// threads can share one pair: distances (g[0..3]) and log ratios (g[4..6]).
__device__ __forceinline__ void pair_geometry_dist(float4 cb, float4 nb, float iou, float mult,
                                                   float* g) {
  const float c_w = __fsub_rn(cb.z, cb.x), c_h = __fsub_rn(cb.w, cb.y);
  const float n_w = __fsub_rn(nb.z, nb.x), n_h = __fsub_rn(nb.w, nb.y);
  const float c_scale = __fdiv_rn(__fadd_rn(c_w, c_h), 2.0f);
  const float c_cx = __fadd_rn(cb.x, __fdiv_rn(c_w, 2.0f));
  const float c_cy = __fadd_rn(cb.y, __fdiv_rn(c_h, 2.0f));
  const float n_cx = __fadd_rn(nb.x, __fdiv_rn(n_w, 2.0f));
  const float n_cy = __fadd_rn(nb.y, __fdiv_rn(n_h, 2.0f));
  const float dx = __fsub_rn(n_cx, c_cx), dy = __fsub_rn(n_cy, c_cy);
  const float l2 = __fdiv_rn(__fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy))), c_scale);
  g[0] = __fmul_rn(iou, mult);
  g[1] = __fmul_rn(__fdiv_rn(dx, c_scale), mult);
  g[2] = __fmul_rn(__fdiv_rn(dy, c_scale), mult);
  g[3] = __fmul_rn(l2, mult);
}
__device__ __forceinline__ void pair_geometry_logs(float4 cb, float4 nb, float mult, float* g) {
  const float c_w = __fsub_rn(cb.z, cb.x), c_h = __fsub_rn(cb.w, cb.y);
  const float n_w = __fsub_rn(nb.z, nb.x), n_h = __fsub_rn(nb.w, nb.y);
  const float w_diff = __fdiv_rn(logf(__fdiv_rn(n_w, c_w)), GN_LN2_F32);
  const float h_diff = __fdiv_rn(logf(__fdiv_rn(n_h, c_h)), GN_LN2_F32);
  const float asp = __fdiv_rn(
      __fsub_rn(logf(__fdiv_rn(n_w, n_h)), logf(__fdiv_rn(c_w, c_h))), GN_LN2_F32);
  g[0] = __fmul_rn(w_diff, mult);
  g[1] = __fmul_rn(h_diff, mult);
  g[2] = __fmul_rn(asp, mult);
}

}  // namespace gn
