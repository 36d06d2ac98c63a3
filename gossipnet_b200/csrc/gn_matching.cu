// DetectionMatching (A9) and the loss (A10) on the GPU.
//
// Matching: one CTA per image.
//   1. visiting order: detections by DESCENDING score.  All threads compute each
//      detection's rank by counting (score desc, index desc), which is the
//      reference's order whenever scores are distinct; if any two scores of the
//      image are equal, thread 0 re-derives the order with the exact libstdc++
//      introsort decision sequence + reverse (det_matching.cc:95-96), so ties are
//      broken like a g++ build of the reference breaks them.
//   2. GT order: introsort on the ignore flags (det_matching.cc:98), thread 0
//      (G is small; the order among crowd GTs decides which one a detection
//      gets, so it has to be the same permutation).
//   3. greedy scan (det_matching.cc:125-159), warp 0: 32 ranks at a time each
//      lane tests whether ITS detection has any IoU >= 0.5 at all (most do
//      not, and an unmatched detection changes no state), then the candidates
//      are visited strictly in order with the GT scan spread over the lanes:
//      regular GTs -> arg max of (iou, position) among unmatched ones >= 0.5;
//      only if none: first crowd GT >= 0.5.  Equivalent to the sequential loop
//      for non-NaN IoUs.
#include "gn_common.cuh"
#include "gn_introsort.cuh"

namespace gn {

constexpr int DM_THREADS = 256;

struct LessScore {
  const float* k;
  __device__ bool operator()(int32_t i, int32_t j) const { return k[i] < k[j]; }
};
struct LessFlag {
  const uint8_t* k;
  __device__ bool operator()(int32_t i, int32_t j) const { return (k[i] != 0) < (k[j] != 0); }
};

__global__ void __launch_bounds__(DM_THREADS)
detection_matching_kernel(const float* __restrict__ iou, const int64_t* __restrict__ iou_off,
                          const float* __restrict__ score, const uint8_t* __restrict__ ignore,
                          const int32_t* __restrict__ img_off, const int32_t* __restrict__ gt_off,
                          float* __restrict__ labels, float* __restrict__ weights,
                          int32_t* __restrict__ assignment, int32_t* __restrict__ order_ws) {
  extern __shared__ int32_t dm_smem[];
  const int img = blockIdx.x;
  const int d0 = img_off[img], n = img_off[img + 1] - d0;
  const int g0 = gt_off[img], G = gt_off[img + 1] - g0;
  const float* sc = score + d0;
  const uint8_t* ign = ignore + g0;
  const float* M = iou + iou_off[img];
  int32_t* order = order_ws + d0;
  int32_t* gt_order = dm_smem;          // [G]
  int32_t* taken = dm_smem + G;         // [G]
  __shared__ int has_ties;
  const int t = threadIdx.x;

  if (t == 0) has_ties = 0;
  for (int i = t; i < n; i += DM_THREADS) {
    labels[d0 + i] = 0.f;
    weights[d0 + i] = 1.f;
    assignment[d0 + i] = -1;
  }
  for (int k = t; k < G; k += DM_THREADS) {
    gt_order[k] = k;
    taken[k] = 0;
  }
  __syncthreads();

  // ---- 1. visiting order ------------------------------------------------------
  for (int i = t; i < n; i += DM_THREADS) {
    const float si = sc[i];
    int rank = 0, ties = 0;
    for (int j = 0; j < n; ++j) {
      const float sj = __ldg(sc + j);
      rank += (sj > si) || (sj == si && j > i);
      ties += (sj == si);
    }
    order[rank < n ? rank : n - 1] = i;
    if (ties > 1) has_ties = 1;
  }
  if (t == 0 && G > 1) {
    IntroSorter<LessFlag> s{gt_order, LessFlag{ign}};
    s.sort(G);
  }
  __syncthreads();
  if (has_ties) {
    for (int i = t; i < n; i += DM_THREADS) order[i] = i;
    __syncthreads();
    if (t == 0) {
      IntroSorter<LessScore> s{order, LessScore{sc}};
      s.sort(n);
      for (int a = 0, b = n - 1; a < b; ++a, --b) {
        const int32_t tmp = order[a]; order[a] = order[b]; order[b] = tmp;
      }
    }
    __syncthreads();
  }
  if (t >= 32 || G == 0) return;

  // number of regular GTs = first crowd position in gt_order
  int R = 0;
  for (int k = t; k < G; k += 32) R += (ign[gt_order[k]] == 0);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) R += __shfl_xor_sync(0xffffffffu, R, d);

  // ---- 3. greedy ---------------------------------------------------------------
  const float thresh = 0.5f;  // det_matching.cc:73
  for (int base = 0; base < n; base += 32) {
    const int my_rank = base + t;
    const int my_det = my_rank < n ? order[my_rank] : -1;
    bool cand = false;
    if (my_det >= 0) {
      const float* row = M + (size_t)my_det * G;
      for (int k = 0; k < G; ++k) cand |= !(__ldg(row + k) < thresh);
    }
    unsigned todo = __ballot_sync(0xffffffffu, cand);
    while (todo) {
      const int src = __ffs(todo) - 1;
      todo &= todo - 1;
      const int det = __shfl_sync(0xffffffffu, my_det, src);
      const float* row = M + (size_t)det * G;
      // regular GTs: best = (iou, position) lexicographic max among unmatched >= thresh
      float best = -1.f;
      int best_pos = -1;
      for (int k = t; k < R; k += 32) {
        const int gt = gt_order[k];
        const float v = __ldg(row + gt);
        if (!taken[gt] && !(v < thresh) && (v > best || (v == best && k > best_pos))) {
          best = v;
          best_pos = k;
        }
      }
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, d);
        const int op = __shfl_xor_sync(0xffffffffu, best_pos, d);
        if (op >= 0 && (best_pos < 0 || ov > best || (ov == best && op > best_pos))) {
          best = ov;
          best_pos = op;
        }
      }
      int match_pos = best_pos;
      if (match_pos < 0) {
        // no regular GT: the first crowd GT (in gt_order) with IoA >= thresh wins
        for (int k0 = R; k0 < G && match_pos < 0; k0 += 32) {
          const int k = k0 + t;
          const bool ok = k < G && !(__ldg(row + gt_order[k]) < thresh);
          const unsigned m = __ballot_sync(0xffffffffu, ok);
          if (m) match_pos = k0 + __ffs(m) - 1;
        }
      }
      if (match_pos >= 0) {
        const int gt = gt_order[match_pos];
        if (t == 0) {
          taken[gt] = 1;
          labels[d0 + det] = 1.f;
          assignment[d0 + det] = gt;
          if (ign[gt]) weights[d0 + det] = 0.f;
        }
        __syncwarp();
      }
    }
  }
}

// ---------------------------------------------------------------------------
// loss: one CTA per image
// ---------------------------------------------------------------------------
constexpr int LOSS_THREADS = 256;

__global__ void __launch_bounds__(LOSS_THREADS)
loss_fwd_kernel(const float* __restrict__ prediction, const float* __restrict__ labels,
                float* __restrict__ weights_io, const int32_t* __restrict__ assignment,
                const uint8_t* __restrict__ gt_crowd, const int32_t* __restrict__ gt_classes,
                const int32_t* __restrict__ img_off, const int32_t* __restrict__ gt_off,
                const float* __restrict__ class_weights, int normalize, float loss_multiplier,
                float* __restrict__ loss_out, float* __restrict__ dlogit) {
  __shared__ float red[LOSS_THREADS / 32];
  const int img = blockIdx.x;
  const int d0 = img_off[img], n = img_off[img + 1] - d0;
  const int g0 = gt_off[img], G = gt_off[img + 1] - g0;
  const int t = threadIdx.x;
  // d loss / d x_i = scale * w_i * (sigmoid(x_i) - z_i)
  const float scale = loss_multiplier * (normalize ? 1.0f / (float)max(n, 1) : 1.0f);
  float sum = 0.f;
  for (int i = t; i < n; i += LOSS_THREADS) {
    const int a = assignment[d0 + i];
    int det_class = 0;
    if (G > 0) {  // network.py:286-297
      const int idx = max(a, 0);
      const bool crowd = gt_crowd[g0 + idx] != 0;
      if (a >= 0 && !crowd) det_class = gt_classes[g0 + idx];
    }
    const float w = weights_io[d0 + i] * __ldg(class_weights + det_class);
    weights_io[d0 + i] = w;
    const float x = prediction[d0 + i], z = labels[d0 + i];
    // tf.nn.sigmoid_cross_entropy_with_logits: max(x,0) - x z + log1p(exp(-|x|))
    const float e = expf(-fabsf(x));
    const float ce = fmaxf(x, 0.f) - x * z + log1pf(e);
    sum += ce * w;
    if (dlogit != nullptr) {
      const float sig = x >= 0.f ? 1.0f / (1.0f + e) : e / (1.0f + e);
      dlogit[d0 + i] = scale * w * (sig - z);
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
  if ((t & 31) == 0) red[t >> 5] = sum;
  __syncthreads();
  if (t == 0) {
    float tot = 0.f;
    for (int i = 0; i < LOSS_THREADS / 32; ++i) tot += red[i];
    const float normed = n > 0 ? tot / (float)n : nanf("");
    loss_out[3 * img + 0] = tot;
    loss_out[3 * img + 1] = normed;
    loss_out[3 * img + 2] = (normalize ? normed : tot) * loss_multiplier;
  }
}

}  // namespace gn

extern "C" int gn_detection_matching(const float* iou, const int64_t* iou_off,
                                     const float* score, const uint8_t* ignore,
                                     const int32_t* img_off, const int32_t* gt_off,
                                     int num_images, int num_dets, int max_gt,
                                     float* labels, float* weights, int32_t* assignment,
                                     int32_t* workspace, gn_stream_t stream) {
  GN_REQUIRE(num_images >= 0 && num_dets >= 0, "gn_detection_matching: negative size");
  if (num_images == 0) return GN_OK;
  GN_REQUIRE(iou_off && img_off && gt_off && labels && weights && assignment && workspace,
             "gn_detection_matching: null pointer");
  GN_REQUIRE(num_dets == 0 || (score != nullptr), "gn_detection_matching: null score");
  // dynamic shared memory: 2 ints per GT of the largest image
  GN_REQUIRE(max_gt >= 0 && max_gt <= 24 * 1024,
             "gn_detection_matching: max_gt=%d outside [0, 24576]", max_gt);
  const int smem = 2 * (int)sizeof(int32_t) * (max_gt > 0 ? max_gt : 1);
  cudaError_t e = cudaFuncSetAttribute(gn::detection_matching_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) {
    gn::set_error("gn_detection_matching: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    return GN_ERR_CUDA;
  }
  gn::detection_matching_kernel<<<num_images, gn::DM_THREADS, smem, (cudaStream_t)stream>>>(
      iou, iou_off, score, ignore, img_off, gt_off, labels, weights, assignment, workspace);
  GN_CHECK_LAUNCH("gn_detection_matching");
  return GN_OK;
}

extern "C" int gn_loss_fwd(const float* prediction, const float* labels, float* weights_io,
                           const int32_t* assignment, const uint8_t* gt_crowd,
                           const int32_t* gt_classes, const int32_t* img_off,
                           const int32_t* gt_off, int num_images, int num_dets,
                           const float* class_weights, int normalize, float loss_multiplier,
                           float* loss_out, float* dlogit, gn_stream_t stream) {
  GN_REQUIRE(num_images >= 0 && num_dets >= 0, "gn_loss_fwd: negative size");
  if (num_images == 0) return GN_OK;
  GN_REQUIRE(img_off && gt_off && class_weights && loss_out, "gn_loss_fwd: null pointer");
  GN_REQUIRE(num_dets == 0 || (prediction && labels && weights_io && assignment),
             "gn_loss_fwd: null pointer");
  gn::loss_fwd_kernel<<<num_images, gn::LOSS_THREADS, 0, (cudaStream_t)stream>>>(
      prediction, labels, weights_io, assignment, gt_crowd, gt_classes, img_off, gt_off,
      class_weights, normalize, loss_multiplier, loss_out, dlogit);
  GN_CHECK_LAUNCH("gn_loss_fwd");
  return GN_OK;
}
