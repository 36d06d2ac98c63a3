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
//   3. greedy scan (det_matching.cc:125-159).  The sequential loop takes, for a detection,
//      the arg max of (iou, position) among the still unmatched regular GTs with
//      iou >= 0.5, and only if there is none the first crowd GT with iou >= 0.5
//      (equivalent for non-NaN IoUs).  Which GTs are ELIGIBLE does not depend on the order,
//      only which are still unmatched does.  So all threads first build, per detection, the
//      list of its eligible regular GTs sorted by (iou, position) descending (up to DM_K of
//      them; a detection rarely overlaps more than two) and its first eligible crowd GT;
//      the sequential part is then one thread walking short lists in shared memory - a few
//      instructions per detection instead of a warp-wide scan + reduction (1.45 ms -> ~0.1 ms
//      for 8 images of 1000 detections).
#include "gn_common.cuh"
#include "gn_introsort.cuh"

namespace gn {

constexpr int DM_THREADS = 256;
constexpr int DM_K = 6;              // eligible regular GTs kept per detection
constexpr int DM_REC = DM_K + 2;     // + number of eligible regular GTs, + first crowd position

struct LessScore {
  const float* k;
  __device__ bool operator()(int32_t i, int32_t j) const { return k[i] < k[j]; }
};
struct LessFlag {
  const uint8_t* k;
  __device__ bool operator()(int32_t i, int32_t j) const { return (k[i] != 0) < (k[j] != 0); }
};

#ifdef DM_TRACE
__device__ long long dm_trace[8];
#define DM_TR(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) dm_trace[i] = clock64(); } while (0)
#else
#define DM_TR(i) do { } while (0)
#endif

__global__ void __launch_bounds__(DM_THREADS)
detection_matching_kernel(const float* __restrict__ iou, const int64_t* __restrict__ iou_off,
                          const float* __restrict__ score, const uint8_t* __restrict__ ignore,
                          const int32_t* __restrict__ img_off, const int32_t* __restrict__ gt_off,
                          float* __restrict__ labels, float* __restrict__ weights,
                          int32_t* __restrict__ assignment, int32_t* __restrict__ order_ws,
                          int32_t* __restrict__ lists) {
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

  DM_TR(0);
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

  DM_TR(1);
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
  DM_TR(2);
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
  DM_TR(3);
  if (G == 0) return;

  // number of regular GTs = first crowd position in gt_order (block-uniform)
  int R = 0;
  for (int k = 0; k < G; ++k) R += (ign[gt_order[k]] == 0);

  // ---- 3a. per-detection candidate records (all threads) -------------------------
  const float thresh = 0.5f;  // det_matching.cc:73
  int32_t* rec = lists + (size_t)d0 * DM_REC;
  for (int i = t; i < n; i += DM_THREADS) {
    const float* row = M + (size_t)i * G;
    float bv[DM_K];
    int bp[DM_K];
#pragma unroll
    for (int j = 0; j < DM_K; ++j) { bv[j] = -1.f; bp[j] = -1; }
    int cnt = 0;
    for (int k = 0; k < R; ++k) {
      const float v = __ldg(row + gt_order[k]);
      if (v >= thresh) {               // (a NaN is never taken by the arg-max scan either)
        ++cnt;
        // insert (v, k) keeping (value, position) descending; later position wins ties
        float cv = v;
        int cp = k;
#pragma unroll
        for (int j = 0; j < DM_K; ++j) {
          const bool better = bp[j] < 0 || cv > bv[j] || (cv == bv[j] && cp > bp[j]);
          if (better) {
            const float tv = bv[j]; const int tp = bp[j];
            bv[j] = cv; bp[j] = cp;
            cv = tv; cp = tp;
          }
        }
      }
    }
    int fc = -1;
    for (int k = R; k < G && fc < 0; ++k)
      if (!(__ldg(row + gt_order[k]) < thresh)) fc = k;
#pragma unroll
    for (int j = 0; j < DM_K; ++j) rec[(size_t)i * DM_REC + j] = bp[j];
    rec[(size_t)i * DM_REC + DM_K] = cnt;
    rec[(size_t)i * DM_REC + DM_K + 1] = fc;
  }
  __syncthreads();
  DM_TR(4);
  if (t >= 32) return;

  // ---- 3b. greedy, in visiting order: warp 0 stages 32 records, lane 0 walks them --------
  __shared__ int32_t chunk[32][DM_REC + 1];
  for (int base = 0; base < n; base += 32) {
    const int my_rank = base + t;
    const int my_det = my_rank < n ? order[my_rank] : -1;
    chunk[t][DM_REC] = my_det;
    if (my_det >= 0) {
#pragma unroll
      for (int j = 0; j < DM_REC; ++j) chunk[t][j] = rec[(size_t)my_det * DM_REC + j];
    }
    __syncwarp();
    if (t == 0) {
      const int m = min(32, n - base);
      for (int e = 0; e < m; ++e) {
        const int det = chunk[e][DM_REC];
        const int cnt = chunk[e][DM_K], fc = chunk[e][DM_K + 1];
        if (cnt == 0 && fc < 0) continue;            // overlaps nothing: changes no state
        int match_pos = -1;
        for (int j = 0; j < DM_K && j < cnt; ++j) {
          const int pos = chunk[e][j];
          if (!taken[gt_order[pos]]) { match_pos = pos; break; }
        }
        if (match_pos < 0 && cnt > DM_K) {
          // more eligible GTs than the record keeps and all kept ones are taken: full scan
          const float* row = M + (size_t)det * G;
          float best = -1.f;
          for (int k = 0; k < R; ++k) {
            const int gt = gt_order[k];
            const float v = __ldg(row + gt);
            if (!taken[gt] && v >= thresh && (v > best || (v == best && k > match_pos))) {
              best = v;
              match_pos = k;
            }
          }
        }
        if (match_pos < 0) match_pos = fc;           // no regular GT: the first eligible crowd GT
        if (match_pos >= 0) {
          const int gt = gt_order[match_pos];
          taken[gt] = 1;
          labels[d0 + det] = 1.f;
          assignment[d0 + det] = gt;
          if (ign[gt]) weights[d0 + det] = 0.f;
        }
      }
    }
    __syncwarp();
  }
  DM_TR(5);
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

extern "C" int64_t gn_detection_matching_workspace_ints(int num_dets) {
  return (int64_t)(num_dets > 0 ? num_dets : 0) * (1 + gn::DM_REC) + 1;
}

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
  // workspace: visiting order [num_dets], then the candidate records [num_dets, DM_REC]
  gn::detection_matching_kernel<<<num_images, gn::DM_THREADS, smem, (cudaStream_t)stream>>>(
      iou, iou_off, score, ignore, img_off, gt_off, labels, weights, assignment, workspace,
      workspace + num_dets);
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

#ifdef DM_TRACE
extern "C" int gn_detection_matching_trace(long long* host_out) {
  return cudaMemcpyFromSymbol(host_out, gn::dm_trace, sizeof(long long) * 8) == cudaSuccess ? 0 : 1;
}
#endif
